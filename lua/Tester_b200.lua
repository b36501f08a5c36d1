--[[ Tester_b200.lua — fast path for fbcoco.Tester_FRCNN:testOne (Tester_FRCNN.lua:54-139) WITHOUT editing that file:
loaded after `require 'Tester_FRCNN'` (fbcoco.lua switch, INTEGRATION.md §3), it wraps the class method. With the default
test options (one localisation pass, no bbox voting, no rbox scores, one score threshold for all classes) the whole of
testOne after getImages — trunk, fused Foveal + ROI pooling, heads, softmax, BBoxNorm + decode, clamp (:75-78), per-class
gather (:106-115) and utils.nms (:117) — is ONE call, mpn_model_detect_nms, instead of detect() plus 80 utils.nms calls
with a host round trip each; any other option set falls through to the reference method, which then runs on
fbcoco.ImageDetect (lua/ImageDetect_b200.lua) and utils.nms / utils.bbox_vote (lua/utils_b200.lua) unchanged.
Returns exactly what the reference returns: img_boxes (tds.hash of K_j x 5 FloatTensors, rows in nms.c's emission order)
and {output, bbox_pred}. `mpn_tester=reference` in the environment switches the wrapper off.
UNTESTED in the build environment (no Torch-7 there). ]]
local ffi = require 'ffi'
local tds = require 'tds'
local mpn = paths.dofile('mpn_ffi.lua')
local C = mpn.C

local Tester = fbcoco.Tester_FRCNN
local ImageDetect = fbcoco.ImageDetect
local testOne_reference = Tester.testOne

local function uniform(t)                                   -- self.thresh: one value for every class?
   local v = t[1]
   for j = 2, t:nElement() do if t[j] ~= v then return nil end end
   return v
end

function Tester:testOne(i)
   local thr = self.thresh and uniform(self.thresh)
   if os.getenv('mpn_tester') == 'reference' or self.num_iter ~= 1 or (opt and (opt.test_bbox_voting or opt.test_use_rbox_scores))
      or (opt and opt.disable_memory_efficient_forward) or not thr or not self.detec._native then
      return testOne_reference(self, i)
   end
   local dataset = self.dataset
   local timer = torch.Timer()
   local boxes = dataset:getROIBoxes(i):float():contiguous()
   local im = dataset:getImage(i)
   local det = self.detec
   local h = ImageDetect._native(det)
   local img, im_scale = ImageDetect._getImage(det, im)
   local R, nc = boxes:size(1), h.num_classes
   local output, bbox_pred = torch.FloatTensor(R, nc), torch.FloatTensor(R, 4 * nc)
   local keep = torch.IntTensor(nc - 1, R)
   local counts = torch.IntTensor(nc - 1)
   local ctx = mpn.ctx()
   -- clamp to the ORIGINAL image (im:size(3) x im:size(2), Tester_FRCNN.lua:75-78), gather with score > thresh, NMS
   mpn.check(ctx, C.mpn_model_detect_nms(h.handle, mpn.fptr(img), img:size(2), img:size(3), mpn.fptr(boxes), R, im_scale,
                                         im:size(3), im:size(2), thr, self.nms_thresh, mpn.fptr(output), mpn.fptr(bbox_pred),
                                         ffi.cast('int32_t*', keep:data()), ffi.cast('int32_t*', counts:data())), 'mpn_model_detect_nms')
   local img_boxes = tds.hash()
   for j = 1, nc - 1 do
      local k = counts[j]
      if k > 0 then
         local idx = keep[j]:narrow(1, 1, k):long():add(1)  -- 0-based proposal rows -> 1-based
         local sb = torch.FloatTensor(k, 5)
         sb:narrow(2, 1, 4):copy(bbox_pred:narrow(2, j * 4 + 1, 4):index(1, idx))
         sb:select(2, 5):copy(output:select(2, j + 1):index(1, idx))
         img_boxes[j] = sb
      else
         img_boxes[j] = torch.FloatTensor()
      end
   end
   print(('test: (%s) %5d/%-5d dev: %d, total time: %.3fs (mpn_model_detect_nms)'):format(dataset.dataset_name, i, dataset:size(),
         cutorch and cutorch.getDevice() or 0, timer:time().real))
   return img_boxes, {output, bbox_pred}
end
