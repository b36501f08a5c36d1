/* Minimal stand-in for Torch-7's <TH/TH.h>: ONLY what /root/reference/nms.c uses
 * (nms.c:9 includes it; nms.c:45-50,61-63,102-103,112-126 use these symbols).
 * TEST INFRASTRUCTURE — lets the literal reference nms.c compile without Torch-7.
 * Written from scratch for this repo; it is not Torch source. */
#ifndef MPN_TH_STUB_H
#define MPN_TH_STUB_H
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef struct THFloatStorage { float *data; long size; } THFloatStorage;
typedef struct THFloatTensor {
  long *size; long *stride; int nDimension;
  THFloatStorage *storage; long storageOffset;
} THFloatTensor;

#define THAssert(c) do { if (!(c)) { fprintf(stderr, "THAssert failed: %s (%s:%d)\n", #c, __FILE__, __LINE__); abort(); } } while (0)

static inline float *THFloatTensor_data(const THFloatTensor *t) {
  return t->storage ? t->storage->data + t->storageOffset : NULL;
}
static inline void th_stub_resize(THFloatTensor *t, int nd, long d0, long d1) {
  long n = d0 * (nd > 1 ? d1 : 1);
  if (!t->storage) t->storage = (THFloatStorage *)calloc(1, sizeof(THFloatStorage));
  if (t->storage->size < n) {
    t->storage->data = (float *)realloc(t->storage->data, sizeof(float) * (n > 0 ? n : 1));
    t->storage->size = n;
  }
  t->size = (long *)realloc(t->size, sizeof(long) * 2);
  t->stride = (long *)realloc(t->stride, sizeof(long) * 2);
  t->nDimension = nd; t->storageOffset = 0;
  t->size[0] = d0; t->size[1] = (nd > 1 ? d1 : 1);
  t->stride[0] = (nd > 1 ? d1 : 1); t->stride[1] = 1;
}
static inline void THFloatTensor_resize1d(THFloatTensor *t, long d0) { th_stub_resize(t, 1, d0, 1); }
static inline void THFloatTensor_resize2d(THFloatTensor *t, long d0, long d1) { th_stub_resize(t, 2, d0, d1); }
static inline void THFloatTensor_resizeAs(THFloatTensor *t, THFloatTensor *s) {
  th_stub_resize(t, s->nDimension, s->size[0], s->nDimension > 1 ? s->size[1] : 1);
}
static inline void THFloatTensor_zero(THFloatTensor *t) {
  long n = t->size[0] * (t->nDimension > 1 ? t->size[1] : 1);
  memset(THFloatTensor_data(t), 0, sizeof(float) * n);
}
static inline int THFloatTensor_isContiguous(const THFloatTensor *t) {
  if (t->nDimension == 1) return t->stride[0] == 1;
  return t->stride[1] == 1 && t->stride[0] == t->size[1];
}
#endif
