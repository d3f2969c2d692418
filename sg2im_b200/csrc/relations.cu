// COCO relation synthesis for a whole batch on the device.
//
// The reference builds the scene graph of every COCO sample on the host inside
// CocoSceneGraphDataset.__getitem__ (sg2im/data/coco.py:294-356): per object a masked centroid
// (two linspace grids, a boolean gather and two means: ~10 small tensor ops per object), then per
// real object one random partner, a random direction and a geometric predicate from the two
// boxes / centroids, then one __in_image__ triple per real object; coco_collate_fn
// (coco.py:376-419) offsets the indices.  Here one launch computes all centroids (a warp per
// object) and one launch writes the collated (T, 3) triple table + triple_to_img.  The random
// draws stay on the host (Python's `random`, same call order as the reference, see
// batching.coco_relation_draws) and come in as two small arrays, so a seeded run reproduces the
// reference's graphs.
//
// Integer outputs; the only floating-point work is the centroid (fp32, a few hundred adds).  The
// angle test of the reference (atan2 against +-pi/4, +-3pi/4 in double) is evaluated as sign /
// magnitude comparisons of the fp32 centroid difference, which selects the same sector for every
// fp32 input (|dy| == |dx| falls on the closed side of each interval exactly as atan2 does; any
// other fp32 pair is at least 2^-24 away from the diagonal, far outside double rounding).
#include "common.cuh"

namespace {

// torch.linspace(a, b, n)[j] in fp32 (ATen scalar formula: ascending from a below the midpoint,
// descending from b above it)
__device__ __forceinline__ float lin(float a, float b, float step, int j, int n) {
  return j < n / 2 ? a + step * (float)j : b - step * (float)(n - 1 - j);
}

// one warp per object
__global__ void __launch_bounds__(256)
rel_centers_kernel(const float* __restrict__ boxes, const int64_t* __restrict__ masks, int MH, int MW,
                   int64_t O, float* __restrict__ centers) {
  const int64_t o = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (o >= O) return;                                   // whole warp leaves together
  const float x0 = boxes[4 * o], y0 = boxes[4 * o + 1], x1 = boxes[4 * o + 2], y1 = boxes[4 * o + 3];
  const float sx = MW > 1 ? (x1 - x0) / (float)(MW - 1) : 0.f;
  const float sy = MH > 1 ? (y1 - y0) / (float)(MH - 1) : 0.f;
  const int64_t* m = masks + o * MH * MW;
  float ax = 0.f, ay = 0.f;
  int cnt = 0;
  for (int i = lane; i < MH * MW; i += 32) {
    if (m[i] == 1) {
      int r = i / MW, c = i - r * MW;
      ax += MW > 1 ? lin(x0, x1, sx, c, MW) : x0;
      ay += MH > 1 ? lin(y0, y1, sy, r, MH) : y0;
      ++cnt;
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    ax += __shfl_xor_sync(0xffffffffu, ax, d);
    ay += __shfl_xor_sync(0xffffffffu, ay, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
  }
  if (lane == 0) {
    // empty mask: the box centre (coco.py:305-307)
    centers[2 * o] = cnt ? ax / (float)cnt : 0.5f * (x0 + x1);
    centers[2 * o + 1] = cnt ? ay / (float)cnt : 0.5f * (y0 + y1);
  }
}

struct PredIds { int64_t left, right, above, below, inside, surrounding, in_image; };

// one thread per object (objects grouped by image, __image__ last in every image)
__global__ void __launch_bounds__(256)
rel_triples_kernel(const float* __restrict__ boxes, const float* __restrict__ centers,
                   const int64_t* __restrict__ obj_off, const int64_t* __restrict__ trip_off,
                   const int64_t* __restrict__ obj_to_img, const int64_t* __restrict__ partner,
                   const uint8_t* __restrict__ swap, int64_t O, PredIds ids,
                   int64_t* __restrict__ triples, int64_t* __restrict__ triple_to_img) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= O) return;
  const int64_t n = obj_to_img[i];
  const int64_t first = obj_off[n], last = obj_off[n + 1] - 1;      // `last` is the __image__ object
  if (i == last) return;
  const int64_t n_real = last - first;
  const int64_t t0 = trip_off[n];
  const int64_t n_rel = trip_off[n + 1] - t0 - n_real;              // n_real, or 0 (lone object / switched off)
  if (n_rel > 0) {
    const int64_t other = partner[i];
    const int64_t s = swap[i] ? other : i, o = swap[i] ? i : other;
    const float sx0 = boxes[4 * s], sy0 = boxes[4 * s + 1], sx1 = boxes[4 * s + 2], sy1 = boxes[4 * s + 3];
    const float ox0 = boxes[4 * o], oy0 = boxes[4 * o + 1], ox1 = boxes[4 * o + 2], oy1 = boxes[4 * o + 3];
    const float dx = centers[2 * s] - centers[2 * o];
    const float dy = centers[2 * s + 1] - centers[2 * o + 1];
    int64_t p;
    if (sx0 < ox0 && sx1 > ox1 && sy0 < oy0 && sy1 > oy1) p = ids.surrounding;
    else if (sx0 > ox0 && sx1 < ox1 && sy0 > oy0 && sy1 < oy1) p = ids.inside;
    else if ((dx < 0.f && fabsf(dy) <= -dx) || (dx == 0.f && dy == 0.f && copysignf(1.f, dx) < 0.f))
      p = ids.left;                                                  // |theta| >= 3pi/4 (atan2(+-0, -0) = +-pi)
    else if (dy < 0.f && (dx < 0.f ? -dy > -dx : -dy > dx)) p = ids.above;   // -3pi/4 <= theta < -pi/4
    else if (dy > 0.f && (dx < 0.f ? dy > -dx : dy >= dx)) p = ids.below;    //  pi/4 <= theta < 3pi/4
    else p = ids.right;                                              // -pi/4 <= theta < pi/4 (incl. d = 0)
    const int64_t t = t0 + (i - first);
    triples[3 * t] = s; triples[3 * t + 1] = p; triples[3 * t + 2] = o;
    triple_to_img[t] = n;
  }
  const int64_t t = t0 + n_rel + (i - first);
  triples[3 * t] = i; triples[3 * t + 1] = ids.in_image; triples[3 * t + 2] = last;
  triple_to_img[t] = n;
}
}  // namespace

extern "C" int sg2im_coco_relations(const float* boxes, const int64_t* masks, int64_t MH, int64_t MW,
                                    const int64_t* obj_off, const int64_t* trip_off,
                                    const int64_t* obj_to_img, const int64_t* partner,
                                    const uint8_t* swap, int64_t O, const int64_t* pred_ids,
                                    float* centers, int64_t* triples, int64_t* triple_to_img,
                                    sg2im_stream_t stream) {
  SG_ARG(boxes && masks && obj_off && trip_off && obj_to_img && partner && swap && pred_ids);
  SG_ARG(centers && triples && triple_to_img);
  SG_ARG(O >= 0 && MH >= 1 && MW >= 1 && MH * MW <= (1 << 20));
  if (O == 0) return 0;
  PredIds ids = {pred_ids[0], pred_ids[1], pred_ids[2], pred_ids[3], pred_ids[4], pred_ids[5], pred_ids[6]};
  cudaStream_t st = as_stream(stream);
  SG_LAUNCH(rel_centers_kernel, (unsigned)ceil_div64(O, 8), 256, 0, st, boxes, masks, (int)MH, (int)MW, O, centers);
  SG_LAUNCH_OK();
  SG_LAUNCH(rel_triples_kernel, (unsigned)ceil_div64(O, 256), 256, 0, st, boxes, centers, obj_off, trip_off,
            obj_to_img, partner, swap, O, ids, triples, triple_to_img);
  SG_LAUNCH_OK();
  return 0;
}
