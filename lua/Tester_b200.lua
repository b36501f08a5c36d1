--[[ Tester_b200.lua — fast path for fbcoco.Tester_FRCNN:testOne (Tester_FRCNN.lua:54-139) WITHOUT editing that file:
loaded after `require 'Tester_FRCNN'` (fbcoco.lua switch, INTEGRATION.md §3), it wraps the class method. With the default
test options (one localisation pass, no bbox voting, no rbox scores, one score threshold for all classes) the whole of
testOne after getImages — trunk, fused Foveal + ROI pooling, heads, softmax, BBoxNorm + decode, clamp (:75-78), per-class
gather (:106-115) and utils.nms (:117) — is ONE call, mpn_model_detect_nms, instead of detect() plus 80 utils.nms calls
with a host round trip each; with iterative localisation, rbox scores or bbox voting it is ONE call as well,
mpn_model_test_one. Per-class thresholds, opt.disable_memory_efficient_forward or more than 8 passes fall through to the
reference method, which then runs on fbcoco.ImageDetect (lua/ImageDetect_b200.lua) and utils.nms / utils.bbox_vote
(lua/utils_b200.lua) unchanged.
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

-- the other option sets (Tester_FRCNN.lua:82-97 iterative localisation, :91-97 rbox scores, :118-124 bbox voting): ONE call as
-- well, mpn_model_test_one — every pass, nn.SelectBoxes between passes on the cached trunk features, the join, the per-class gather,
-- NMS and the voting stay on the device. Returns what the reference returns: img_boxes[j] = the NMS'ed (and voted) K_j x 5 rows,
-- {output, bbox_pred} = the joined outputs (:99-100).
local function testOne_device_tail(self, i, thr)
   local dataset = self.dataset
   local timer = torch.Timer()
   local boxes = dataset:getROIBoxes(i):float():contiguous()
   local im = dataset:getImage(i)
   local det = self.detec
   local h = ImageDetect._native(det)
   local img, im_scale = ImageDetect._getImage(det, im)
   local R, nc = boxes:size(1), h.num_classes
   local o = ffi.new('mpn_test_opts')
   o.num_iter = self.num_iter
   o.use_rbox_scores = (opt and opt.test_use_rbox_scores) and 1 or 0
   o.bbox_voting = (opt and opt.test_bbox_voting) and 1 or 0
   o.score_thresh = thr
   o.nms_thr = self.nms_thresh
   -- the reference passes self.test_bbox_voting_nms_threshold, a field nobody sets (the constructor stores bbox_vote_thresh,
   -- Tester_FRCNN.lua:29 vs :123): the intended 0.5 is kept (SURVEY 8f-2)
   o.vote_thr = self.test_bbox_voting_nms_threshold or self.bbox_vote_thresh or 0.5
   o.vote_score_pow = (opt and opt.test_bbox_voting_score_pow) or 1
   assert(o.use_rbox_scores == 0 or self.num_iter > 1)                      -- assert(#all_output > 1), :92
   local n_out = R * (self.num_iter - o.use_rbox_scores)
   local output, bbox_pred = torch.FloatTensor(n_out, nc), torch.FloatTensor(n_out, 4 * nc)
   local keep = torch.IntTensor(nc - 1, n_out)
   local counts = torch.IntTensor(nc - 1)
   local voted = o.bbox_voting == 1 and torch.FloatTensor(nc - 1, n_out, 5) or nil
   mpn.check(mpn.ctx(), C.mpn_model_test_one(h.handle, mpn.fptr(img), img:size(2), img:size(3), mpn.fptr(boxes), R, im_scale, im:size(3), im:size(2), o,
                                             mpn.fptr(output), mpn.fptr(bbox_pred), ffi.cast('int32_t*', keep:data()), ffi.cast('int32_t*', counts:data()),
                                             voted and mpn.fptr(voted) or nil), 'mpn_model_test_one')
   local img_boxes = tds.hash()
   for j = 1, nc - 1 do
      local k = counts[j]
      if k == 0 then
         img_boxes[j] = torch.FloatTensor()
      elseif voted then
         img_boxes[j] = voted[j]:narrow(1, 1, k):clone()                    -- row i = the voted box of the i-th kept row (nms.c:110-142)
      else
         local idx = keep[j]:narrow(1, 1, k):long():add(1)                  -- 0-based rows of the joined outputs -> 1-based
         local sb = torch.FloatTensor(k, 5)
         sb:narrow(2, 1, 4):copy(bbox_pred:narrow(2, j * 4 + 1, 4):index(1, idx))
         sb:select(2, 5):copy(output:select(2, j + 1):index(1, idx))
         img_boxes[j] = sb
      end
   end
   print(('test: (%s) %5d/%-5d dev: %d, total time: %.3fs (mpn_model_test_one, %d passes)'):format(dataset.dataset_name, i, dataset:size(),
         cutorch and cutorch.getDevice() or 0, timer:time().real, self.num_iter))
   return img_boxes, {output, bbox_pred}
end

function Tester:testOne(i)
   local thr = self.thresh and uniform(self.thresh)
   if os.getenv('mpn_tester') == 'reference' or (opt and opt.disable_memory_efficient_forward) or not thr or not self.detec._native then
      return testOne_reference(self, i)
   end
   if self.num_iter ~= 1 or (opt and (opt.test_bbox_voting or opt.test_use_rbox_scores)) then
      -- mpn_model_test_one takes 1..8 passes; a voting score power other than 1 stays with the reference's own pow() (the device
      -- powf is not pinned bit for bit against it: the Python mirror draws the same line, multipathnet_b200/tester.py)
      local pw = opt and opt.test_bbox_voting and (opt.test_bbox_voting_score_pow or 1) or 1
      if self.num_iter > 8 or pw ~= 1 then return testOne_reference(self, i) end
      return testOne_device_tail(self, i, thr)
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
