--[[ utils_b200.lua — drop-in replacements for the native-backed functions of utils.lua.

   local utils = paths.dofile'utils.lua'
   if os.getenv('mpn_backend') == 'b200' then paths.dofile'lua/utils_b200.lua'(utils) end

keeps every caller (Tester_FRCNN.lua:117,123, demo.lua:85) unchanged. UNTESTED here (no LuaJIT). ]]
local ffi = require 'ffi'
local mpn = paths.dofile('mpn_ffi.lua')
local C = mpn.C

return function(utils)
   -- utils.nms(boxes, overlap) -> FloatTensor K x 5 of kept ROWS (utils.lua:29-33, nms.c:59-108)
   function utils.nms(boxes, overlap)
      local n = boxes:nElement() > 0 and boxes:size(1) or 0
      local keep = torch.FloatTensor()
      if n == 0 then return keep end
      local b = boxes:float():contiguous()
      local idx = torch.IntTensor(n)
      local cnt = ffi.new('int64_t[1]')
      local ctx = mpn.ctx()
      mpn.check(ctx, C.mpn_nms(ctx, mpn.fptr(b), n, overlap, ffi.cast('int32_t*', idx:data()), cnt), 'mpn_nms')
      local k = tonumber(cnt[0])
      if k == 0 then return keep end
      return b:index(1, idx:narrow(1, 1, k):long():add(1))      -- 0-based -> 1-based rows
   end

   -- utils.nms_dense(boxes, overlap) -> LongTensor of 1-based indices (utils.lua:402-462)
   function utils.nms_dense(boxes, overlap)
      local n = boxes:nElement() > 0 and boxes:size(1) or 0
      if n == 0 then return torch.LongTensor() end
      assert(boxes:size(2) == 5)
      local b = boxes:float():contiguous()
      local idx = torch.IntTensor(n)
      local cnt = ffi.new('int64_t[1]')
      local ctx = mpn.ctx()
      mpn.check(ctx, C.mpn_nms_dense(ctx, mpn.fptr(b), n, overlap, ffi.cast('int32_t*', idx:data()), cnt), 'mpn_nms_dense')
      return idx:narrow(1, 1, tonumber(cnt[0])):long():add(1)
   end

   -- utils.bbox_vote(nms_boxes, scored_boxes, overlap) (utils.lua:35-39, nms.c:110-142)
   function utils.bbox_vote(nms_boxes, scored_boxes, overlap)
      local res = torch.FloatTensor():resizeAs(nms_boxes):zero()
      if nms_boxes:nElement() == 0 then return res end
      local a, s = nms_boxes:float():contiguous(), scored_boxes:float():contiguous()
      local ctx = mpn.ctx()
      mpn.check(ctx, C.mpn_bbox_vote(ctx, mpn.fptr(a), a:size(1), mpn.fptr(s), s:size(1), overlap, mpn.fptr(res)), 'mpn_bbox_vote')
      return res
   end
   return utils
end
