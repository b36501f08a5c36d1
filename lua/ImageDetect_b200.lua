--[[ ImageDetect_b200.lua — fbcoco.ImageDetect with the same constructor / detect() signature
(ImageDetect.lua:12-20,156-193) whose forward runs in libmpn_b200.so. Tester_FRCNN.lua:24,72,86 and
demo.lua:43,75 call it unchanged. The model argument is an mpn model description table produced by
lua/model_desc.lua from the nn graph (weights copied once). getImages runs in Lua exactly as in the reference
by default; with `mpn_getimages=device` in the environment the raw image goes to mpn_model_trunk_image, which
applies the transformer and image.scale on the GPU (SURVEY 8f-1) and keeps the features for mpn_model_detect.
UNTESTED in the build environment (no Torch-7 there). ]]
local ffi = require 'ffi'
local mpn = paths.dofile('mpn_ffi.lua')
local C = mpn.C
local ImageDetect = torch.class('fbcoco.ImageDetect')

function ImageDetect:__init(model, transformer, scale, max_size)
   assert(model, 'must provide model!')
   assert(transformer, 'must provide transformer!')
   self.model = model            -- table {handle = mpn_model*, num_classes = C}; see lua/model_desc.lua
   self.image_transformer = transformer
   self.scale = scale or {600}
   self.max_size = max_size or 1000
end

local function getImage(self, im)   -- ImageDetect.lua:22-52, single scale
   im = self.image_transformer:forward(im)
   local s = im[1]:size()
   local smin, smax = math.min(s[1], s[2]), math.max(s[1], s[2])
   local im_scale = self.scale[1] / smin
   if torch.round(im_scale * smax) > self.max_size then im_scale = self.max_size / smax end
   return image.scale(im, s[2] * im_scale, s[1] * im_scale):float():contiguous(), im_scale
end

local on_device = os.getenv('mpn_getimages') == 'device'
local function transform_struct(t)   -- fbcoco.ImageTransformer fields (ImageTransformer.lua:11-17) -> mpn_image_transform
   local tf = ffi.new('mpn_image_transform')
   for i = 1, 3 do
      tf.swap[i - 1] = t.swap and t.swap[i] or i
      tf.mean[i - 1] = t.mean[i]
      tf.std[i - 1] = t.std and t.std[i] or 1
   end
   tf.scale = t.scale or 1
   tf.has_std = t.std and 1 or 0
   return tf
end

-- supposes boxes is in [x1,y1,x2,y2] format
function ImageDetect:detect(im, boxes, min_images, recompute_features)
   if recompute_features == nil then recompute_features = true end
   local b = boxes:float():contiguous()
   local R, nc = b:size(1), self.model.num_classes
   local img, im_scale
   if recompute_features and on_device then
      local raw = im:float():contiguous()
      local s, h, w = ffi.new('double[1]'), ffi.new('int32_t[1]'), ffi.new('int32_t[1]')
      mpn.check(mpn.ctx(), C.mpn_model_trunk_image(self.model.handle, mpn.fptr(raw), raw:size(2), raw:size(3),
                transform_struct(self.image_transformer), self.scale[1], self.max_size, s, h, w), 'mpn_model_trunk_image')
      im_scale = s[0]; self._im_scale = im_scale
      recompute_features = false                      -- the trunk has run: heads on the cached features
   elseif recompute_features then
      img, im_scale = getImage(self, im); self._im_scale = im_scale
   else
      im_scale = self._im_scale
   end
   local scores, bboxes = torch.FloatTensor(R, nc), torch.FloatTensor(R, 4 * nc)
   local m = self.model.handle
   local rc = C.mpn_model_detect(m, img and mpn.fptr(img) or nil, img and img:size(2) or 0, img and img:size(3) or 0,
                                 mpn.fptr(b), R, im_scale, recompute_features and 1 or 0, mpn.fptr(scores), mpn.fptr(bboxes))
   mpn.check(mpn.ctx(), rc, 'mpn_model_detect')
   return scores, bboxes
end
