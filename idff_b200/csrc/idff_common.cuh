// Shared declarations of libidff_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>

#include "../../include/idff_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "idff_b200 is written for sm_100a (B200) only"
#endif

namespace idff {

// ---- error plumbing ---------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define IDFF_CHECK_ARG(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::idff::set_error(__VA_ARGS__);    \
      return IDFF_EINVAL;                \
    }                                    \
  } while (0)

#define IDFF_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      ::idff::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      return IDFF_ECUDA;                                                                             \
    }                                                                                                \
  } while (0)

// Entry points without a handle launch on the device that owns their buffers, whatever device is current on the calling
// thread (a tensor on cuda:1 while cuda:0 is current must not launch on cuda:0).
struct PtrDeviceGuard {
  int prev = -1, dev = 0;
  explicit PtrDeviceGuard(const void* p) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    dev = prev < 0 ? 0 : prev;
    cudaPointerAttributes a;
    if (p != nullptr && cudaPointerGetAttributes(&a, p) == cudaSuccess && a.type == cudaMemoryTypeDevice && a.device != dev &&
        cudaSetDevice(a.device) == cudaSuccess)
      dev = a.device;
    else
      cudaGetLastError();   // a host pointer in a validation test is not an error here
  }
  ~PtrDeviceGuard() {
    if (prev >= 0 && dev != prev) cudaSetDevice(prev);
  }
};

// ---- the packed network --------------------------------------------------------------------------------------
// Device view of one 4-layer SELU MLP (torchcfm/models/models.py:4-21).  All pointers are device pointers.
struct NetView {
  int din;       // layer-0 input columns that depend on the particle: [x.., t] (embedding columns folded away)
  int din_pad;   // din rounded up to 4
  int w;         // hidden width
  int dout;      // network outputs
  int n_tidx;    // rows of the folded layer-0 bias table (1 without an embedding)
  const float* wt0;   // [din_pad][w]   W0^T (rows >= din are zero)
  const float* wt1;   // [w][w]         W1^T : forward operand  (row k = input unit k, contiguous over outputs)
  const float* wt2;   // [w][w]         W2^T
  const float* w1;    // [w][w]         W1 as stored by nn.Linear ([out][in]): reverse-sweep operand
  const float* w2;    // [w][w]         W2
  const float* w3;    // [dout][w]      W3 as stored
  const float* wt3;   // [w][dout]      W3^T
  const float* b0;    // [n_tidx][w]    layer-0 bias (+ W0[:, din:] . emb[t_idx])
  const float* b1;    // [w]
  const float* b2;    // [w]
  const float* b3;    // [dout]
  const float* u3;    // [w]  sum_k W3[k][n]   : adjoint seed of phi = sum_k out_k
  const float* zd1;   // [w]  sum_{d<D} W0[n][d] : tangent of z1 along d = (1,..,1)  (D = din - 1)
  // column-permuted copies for the fused kernel (w in {128, 256}; null otherwise).  Physical column p of a row
  // holds logical unit perm(p) = l + 32*(4h + jj) with p = 128h + 4l + jj, so that lane l of a warp reads its
  // units {l, l+32, ...} with conflict-free 128-bit loads and writes them back with unit-stride addresses.
  const float* fwt1;  // [w][w]  fwt1[k][p] = W1[perm(p)][k]   forward,  layer 2
  const float* fwt2;  // [w][w]  fwt2[k][p] = W2[perm(p)][k]   forward,  layer 3
  const float* fw2r;  // [w][w]  fw2r[n][p] = W2[n][perm(p)]   reverse,  layer 3 -> 2
  const float* fw1r;  // [w][w]  fw1r[n][p] = W1[n][perm(p)]   reverse,  layer 2 -> 1
  const float* w0n;   // [w][din-1]  W0[n][d], d < D            reverse,  layer 1 -> x  (wide-D path)
  // tcgen05 3xTF32 operand tiles of k_tc (w == 256, 3 -> w -> w -> w -> 2 networks): [gemm 0..3][m-tile < w/128][k-block < w/32][hi, lo]
  // [128 rows x 32 tf32], each tile a
  // 16 KB shared-memory image in the canonical K-major SWIZZLE_128B layout (row r at r*128 B, 16-byte chunk c stored at
  // chunk c ^ (r & 7)), so one cp.async.bulk lands it ready for the UMMA descriptor.  gemm 0: W1 (rows = out units),
  // 1: W2, 2: W2^T, 3: W1^T.  hi = rn_tf32(w), lo = rn_tf32(w - hi)  (3xTF32 split).
  const float* tc_tiles;
  // 3xFP16 operand tiles of the fused fp16 tensor kernel (w == 256 only): [gemm 0..3][m-tile 0..1][k-block 0..3][hi, lo']
  // [128 rows x 64 fp16], each a 16 KB K-major SWIZZLE_128B image (row r at r*128 B, 16-byte chunk c = 8 halfs at chunk
  // c ^ (r & 7)).  hi = fp16_rn(w), lo' = fp16_rn((w - hi) * 2^11).  Same gemm numbering as tc_tiles.  Tile row r holds
  // unit (r & ~31) + 4*slot(r & 7) + ((r >> 3) & 3) of the m-tile (see idff_sampler_tc16.cu).
  const float* tc16_tiles;
  // 3xFP16 operand tiles of the layered engine (w % 128 == 0, 256..2048): [gemm 0..3][m-tile < w/128][k-block < w/64][hi, lo']
  // [128 rows x 64 fp16], same images as tc16_tiles but with the rows in natural unit order (tile row r = unit r).
  const float* tcw16_tiles;
  // 3xFP16 operand tiles of the PolyALA tensor kernel (w == 128, din - 1 == dout <= 64; idff_md_tc16.cu): 13 tiles [128 rows x 64 k]
  // hi | lo' (32 KB each, rows permuted like tc16_tiles): 0: W0[:, :D] (k = state dimension, zero beyond D); 1-2: W1; 3-4: W2;
  // 5-6: W3 (rows = outputs, zero beyond dout); 7-8: W2^T; 9-10: W1^T; 11-12: W0[:, :D]^T (rows = state dimensions).
  const float* md16_tiles;
};

}  // namespace idff

struct idff_weights {
  idff::NetView v;
  int device;
  int dims[5];
  int emb_rows, emb_dim;
  float* d_blob;        // single allocation holding every array above
  size_t blob_floats;
  float* d_stash;       // scratch owned by the handle (order-3 stash of the fused kernel, buffers of the layered engine)
  size_t stash_bytes;
  int tc_verified;      // bit 0 / 1: the first fused / layered tcgen05 launch on this handle was checked (alignment flag)
  void* d_sweep;        // device copy of the SweepTab array of the last idff_sample_rk4_sweep call
  size_t sweep_bytes;
  float* d_sweep_out;   // staging buffer of a mixed gamma sweep (one class of sets at a time), grow-only
  size_t sweep_out_bytes;
  unsigned long long* d_counters;   // [0] particle-evaluations of the fp16-split kernels whose field value was not finite,
                                    // [1] rows re-solved in fp32 by the range safety net (idff_fixup.cu)
  float* d_fix_x;       // input rows of a launch that runs in place / continues a chunked solve, kept for the safety net; grow-only
  size_t fix_x_bytes;
  void* d_fix_tab;      // stage tables of the safety net for solves the grid-constant table cannot hold; grow-only
  size_t fix_tab_bytes;
};

namespace idff {

// ---- per-evaluation scalars, computed on the host in fp32 in the reference's operation order ------------------
// One Stage = one call of the field at a scalar time t.
struct Stage {
  float t;        // evaluation time
  int kind;       // 0: t == 0 branch, 1: t == 1 branch, 2: interior
  float a;        // sqrt(1 - Gamma*sigma_t^2)            (multiplies (x1hat - x))
  float den;      // (1 - t)
  float c[3];     // folded momentum coefficients of g1,g2,g3 (sign included); SCOREHEAD: c[0] scales st, c[1] is the constant term
  float end_div;  // divisor of the t == 1 branch
};

void make_stage(const idff_field_params_t& p, float t, Stage* out);



// ---- launch tables (passed by value as __grid_constant__ kernel parameters; CUDA 12.1+ allows 32 KB) ---------
constexpr int kMaxIntervals = 96;
struct Rk4Table {
  int n_iv;
  float dt[kMaxIntervals];
  Stage st[4 * kMaxIntervals];
};

// gamma sweeps (idff_sample_rk4_sweep): one table per parameter set, in global memory
constexpr int kMaxSweepIntervals = 32;
struct SweepTab {
  int reverse;   // jets_needed(): whether this set's interior evaluations run the reverse sweep
  int pad[3];
  float dt[kMaxSweepIntervals];
  Stage st[4 * kMaxSweepIntervals];
};

constexpr int kMaxSdeSteps = 48;
constexpr int kMaxSdeEmit = 64;
struct SdeTable {
  int n_steps, n_emit, method;      // method 0 srk, 1 euler
  float h[kMaxSdeSteps];
  float gval[kMaxSdeSteps][4];      // g at t0 + C1[s]*h (state independent)
  Stage st[kMaxSdeSteps][3];        // f evaluations at t0 + C0[s]*h, s = 0,1,2 (alpha[3] == 0)
  int emit_step[kMaxSdeEmit];       // after this step (-1: before the first) ...
  int emit_out[kMaxSdeEmit];        // ... write output row emit_out = w0*prev_y + w1*y
  float emit_w0[kMaxSdeEmit], emit_w1[kMaxSdeEmit];
};

constexpr float kSeluLambda = 1.0507009873554805f;
constexpr float kSeluAlpha = 1.6732632423543772f;
constexpr float kSeluLA = 1.0507009873554805f * 1.6732632423543772f;

#ifdef __CUDACC__
// SELU value and the derivative code.  ATen evaluates ELU as (exp(x) - 1) * (alpha*scale) for x <= 0 and its
// backward as (alpha*scale) * exp(x) (x <= 0) / scale (x > 0); all higher derivatives on the negative branch
// equal E = alpha*scale*exp(x) and vanish on the positive branch.  `code` packs both: code < 0 -> positive
// branch (s' = lambda, E = 0), else s' = E = code.
__device__ __forceinline__ float selu_fwd(float z, float& code) {
  // branch-free (a warp holds units on both sides of the kink): exp of min(z, 0), then select.  !(z <= 0) keeps NaN.
  const float e = expf(fminf(z, 0.0f));
  const bool pos = !(z <= 0.0f);
  code = pos ? -1.0f : kSeluLA * e;
  return pos ? z * kSeluLambda : (e - 1.0f) * kSeluLA;
}
__device__ __forceinline__ float selu_val(float z) {
  return z <= 0.0f ? (expf(z) - 1.0f) * kSeluLA : z * kSeluLambda;
}
__device__ __forceinline__ void selu_decode(float code, float& s1, float& E) {
  s1 = code < 0.0f ? kSeluLambda : code;
  E = code < 0.0f ? 0.0f : code;
}
#endif

}  // namespace idff
