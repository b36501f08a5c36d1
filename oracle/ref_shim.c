/* TEST INFRASTRUCTURE. Plain-pointer entry points around the LITERAL reference
 * nms.c (compiled from /root/reference/nms.c where it lies, never copied):
 * builds THFloatTensor views with the stub TH.h and calls the reference's own
 * NMS / bbox_vote / boxoverlap (nms.c:43,59,110). Linked into oracle/_ref/libnms_ref.so. */
#include <TH/TH.h>
void NMS(THFloatTensor *keep, THFloatTensor *scored_boxes, float threshold);
void bbox_vote(THFloatTensor *res, THFloatTensor *nms_boxes, THFloatTensor *scored_boxes, float threshold);
void boxoverlap(THFloatTensor *result, THFloatTensor *a, THFloatTensor *b);

static void view2d(THFloatTensor *t, THFloatStorage *s, long sz[2], long st[2], float *p, long n, long m) {
  s->data = p; s->size = n * m; sz[0] = n; sz[1] = m; st[0] = m; st[1] = 1;
  t->size = sz; t->stride = st; t->nDimension = 2; t->storage = s; t->storageOffset = 0;
}
static void release(THFloatTensor *t) {
  if (t->storage) { free(t->storage->data); free(t->storage); }
  free(t->size); free(t->stride);
}
/* returns number kept; keep_rows (capacity N*5) receives the kept rows in selection order */
long ref_nms(const float *scored_boxes, long N, float thr, float *keep_rows) {
  THFloatTensor in, keep; THFloatStorage sin; long sz[2], st[2];
  memset(&keep, 0, sizeof(keep));
  view2d(&in, &sin, sz, st, (float *)scored_boxes, N, 5);
  NMS(&keep, &in, thr);
  long k = keep.size[0];
  if (k > 0) memcpy(keep_rows, THFloatTensor_data(&keep), sizeof(float) * 5 * k);
  release(&keep);
  return k;
}
void ref_bbox_vote(const float *nms_boxes, long K, const float *scored_boxes, long N, float thr, float *res) {
  THFloatTensor a, b, r; THFloatStorage sa, sb; long sza[2], sta[2], szb[2], stb[2];
  memset(&r, 0, sizeof(r));
  view2d(&a, &sa, sza, sta, (float *)nms_boxes, K, 5);
  view2d(&b, &sb, szb, stb, (float *)scored_boxes, N, 5);
  bbox_vote(&r, &a, &b, thr);
  if (K > 0) memcpy(res, THFloatTensor_data(&r), sizeof(float) * 5 * K);
  release(&r);
}
void ref_boxoverlap(const float *a_N4, long N, const float *b4, float *out) {
  THFloatTensor a, b, r; THFloatStorage sa, sb; long sza[2], sta[2], szb[2], stb[2];
  memset(&r, 0, sizeof(r));
  view2d(&a, &sa, sza, sta, (float *)a_N4, N, 4);
  view2d(&b, &sb, szb, stb, (float *)b4, 1, 4);
  boxoverlap(&r, &a, &b);
  if (N > 0) memcpy(out, THFloatTensor_data(&r), sizeof(float) * N);
  release(&r);
}
