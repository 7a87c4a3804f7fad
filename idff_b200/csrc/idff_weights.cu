// Weights handle, error plumbing and the host-side per-evaluation scalars.
#include <cmath>
#include <cstdarg>
#include <atomic>
#include <vector>

#include <cuda_fp16.h>

#include "idff_common.cuh"

namespace idff {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Scalars of one field evaluation, in fp32 and in the operation order of the reference wrappers
// (Autograd_example_order2.py:47,73-75; example_order1_with_prediction_head.py:75,93-95;
//  MD_simulation.py:502,523-525).  The per-element products c*g*sigma_t2/sigma^2 are folded into one scalar.
void make_stage(const idff_field_params_t& p, float t, Stage* out) {
  Stage s;
  memset(&s, 0, sizeof(s));
  s.t = t;
  s.kind = (t == 0.0f) ? 0 : (t == 1.0f ? 1 : 2);
  s.end_div = (float)p.end_div;
  volatile float one_m_t = 1.0f - t;
  s.den = one_m_t;
  const float sigf = (float)p.sigma;
  const float sig2f = (float)(p.sigma * p.sigma);
  const float c1f = (float)(0.5 * (2.0 * p.gamma[0] - 1.0));
  if (p.variant == IDFF_FIELD_MOMENTUM) {
    volatile float prod = t * one_m_t;
    volatile float sq = sqrtf(prod);
    volatile float sig_t = sigf * sq;
    volatile float sig_t2 = sig_t * sig_t;
    double G = 0.0;
    for (int i = 0; i < p.order; ++i) G += p.gamma[i];
    volatile float gs = (float)G * sig_t2;
    volatile float om = 1.0f - gs;
    s.a = sqrtf(om);
    volatile float m0 = c1f * sig_t2;
    s.c[0] = m0 / sig2f;
    for (int i = 1; i < p.order && i < 3; ++i) {
      volatile float mi = (float)p.gamma[i] * sig_t2;
      s.c[i] = mi / sig2f;
    }
  } else if (p.variant == IDFF_FIELD_MD) {
    volatile float m = sig2f * t;
    volatile float sig_t2 = m * one_m_t;
    volatile float om = 1.0f - sig_t2;
    s.a = sqrtf(om);
    volatile float m0 = c1f * sig_t2;
    s.c[0] = -(m0 / sig2f);
    volatile float m1 = (float)p.gamma[1] * sig_t2;
    s.c[1] = -(m1 / sig2f);
  } else if (p.variant == IDFF_FIELD_SCOREHEAD) {
    volatile float prod = t * one_m_t;
    volatile float sq = sqrtf(prod);
    volatile float sig_t = sigf * sq;
    volatile float sig_t2 = sig_t * sig_t;
    volatile float gs = (float)(p.gamma[0] + p.gamma[1]) * sig_t2;
    volatile float om = 1.0f - gs;
    s.a = sqrtf(om);
    s.c[0] = c1f * sig_t2;
    s.c[1] = (float)p.gamma[1] * sig_t2;
  } else {
    s.a = 1.0f;
  }
  *out = s;
}

}  // namespace idff

using namespace idff;

#ifndef IDFF_SRC_HASH
#define IDFF_SRC_HASH "unstamped"
#endif
extern "C" int idff_version(void) { return 200; }
extern "C" const char* idff_source_hash(void) { return IDFF_SRC_HASH; }
extern "C" const char* idff_last_error(void) { return idff::g_err; }
extern "C" uint64_t idff_launch_count(void) { return idff::g_launches.load(); }

extern "C" int idff_weights_create(const float* const* h_W, const float* const* h_b, const int32_t* dims,
                                   int32_t n_layers, const float* h_emb, int32_t emb_rows, int32_t emb_dim,
                                   int32_t device, idff_weights_t** out) {
  IDFF_CHECK_ARG(h_W && h_b && dims && out, "idff_weights_create: null argument");
  IDFF_CHECK_ARG(n_layers == 4, "idff_weights_create: n_layers must be 4 (Linear-SELU x3 -Linear), got %d", n_layers);
  const int din_total = dims[0], w = dims[1], dout = dims[4];
  IDFF_CHECK_ARG(dims[2] == w && dims[3] == w, "idff_weights_create: hidden widths differ (%d,%d,%d)", dims[1], dims[2], dims[3]);
  IDFF_CHECK_ARG(w >= 32 && w % 32 == 0 && w <= 4096, "idff_weights_create: hidden width %d must be a multiple of 32 in [32,4096]", w);
  const bool has_emb = h_emb != nullptr;
  IDFF_CHECK_ARG(!has_emb || (emb_rows > 0 && emb_dim > 0 && emb_dim < din_total), "idff_weights_create: bad embedding shape");
  const int din = has_emb ? din_total - emb_dim : din_total;
  IDFF_CHECK_ARG(din >= 1 && din <= 256 && dout >= 1 && dout <= 256, "idff_weights_create: din %d / dout %d out of range", din, dout);
  const int din_pad = (din + 3) & ~3;
  const int n_tidx = has_emb ? emb_rows : 1;

  // layout of the blob (all offsets multiples of 4 floats = 16 B, the TMA bulk-copy granule)
  auto al = [](size_t n) { return (n + 31) & ~size_t(31); };
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off += al(n); return o; };
  const size_t o_wt0 = take((size_t)din_pad * w), o_wt1 = take((size_t)w * w), o_wt2 = take((size_t)w * w);
  const size_t o_w1 = take((size_t)w * w), o_w2 = take((size_t)w * w), o_w3 = take((size_t)dout * w);
  const size_t o_wt3 = take((size_t)w * dout), o_b0 = take((size_t)n_tidx * w), o_b1 = take(w), o_b2 = take(w);
  const size_t o_b3 = take(dout), o_u3 = take(w), o_zd1 = take(w);
  const bool fused_layout = (w == 128 || w == 256);
  const size_t o_fwt1 = fused_layout ? take((size_t)w * w) : 0, o_fwt2 = fused_layout ? take((size_t)w * w) : 0;
  const size_t o_fw2r = fused_layout ? take((size_t)w * w) : 0, o_fw1r = fused_layout ? take((size_t)w * w) : 0;
  const size_t o_w0n = take((size_t)w * (din > 1 ? din - 1 : 1));
  const bool wide_layout = (w % 128 == 0) && w >= 256 && w <= 2048 && din == 3 && dout == 2;   // layered fp16 engine
  const bool tc_layout = wide_layout && w == 256;                                               // k_tc (3xTF32)
  const int tc_mt = w / 128, tc_kb = w / 32;
  const size_t tc_floats = (size_t)4 * tc_mt * tc_kb * 2 * 128 * 32;
  off = (off + 255) & ~size_t(255);   // 1 KB alignment of the tile images
  const size_t o_tc = tc_layout ? take(tc_floats) : 0;
  const int tw_mt = w / 128, tw_kb = w / 64;
  const size_t tw_floats = (size_t)4 * tw_mt * tw_kb * 2 * 4096;   // 16 KB (4096 floats) per hi / lo' image
  off = (off + 255) & ~size_t(255);
  const size_t o_tw = wide_layout ? take(tw_floats) : 0;
  // k_tc16 also serves the score-head network of the order-1 script (5-256-256-256-4: cat[x, x, t] -> (x1hat, st))
  const bool tc16_layout = (tc_layout && w == 256) || (w == 256 && din == 5 && dout == 4 && !has_emb);
  const size_t tc16_floats = (size_t)4 * 2 * 4 * 2 * 128 * 32;   // 4 gemms x 2 m-tiles x 4 k-blocks x (hi, lo') x 16 KB
  off = (off + 255) & ~size_t(255);
  const size_t o_tc16 = tc16_layout ? take(tc16_floats) : 0;
  const bool md16_layout = (w == 128) && din >= 2 && din - 1 == dout && dout <= 64;
  const size_t md16_floats = (size_t)13 * 2 * 4096;
  off = (off + 255) & ~size_t(255);
  const size_t o_md16 = md16_layout ? take(md16_floats) : 0;
  std::vector<float> blob(off, 0.0f);

  const float *W0 = h_W[0], *W1 = h_W[1], *W2 = h_W[2], *W3 = h_W[3];
  for (int n = 0; n < w; ++n)
    for (int k = 0; k < din; ++k) blob[o_wt0 + (size_t)k * w + n] = W0[(size_t)n * din_total + k];
  for (int n = 0; n < w; ++n)
    for (int k = 0; k < w; ++k) {
      blob[o_wt1 + (size_t)k * w + n] = W1[(size_t)n * w + k];
      blob[o_wt2 + (size_t)k * w + n] = W2[(size_t)n * w + k];
    }
  memcpy(&blob[o_w1], W1, sizeof(float) * w * w);
  memcpy(&blob[o_w2], W2, sizeof(float) * w * w);
  memcpy(&blob[o_w3], W3, sizeof(float) * dout * w);
  for (int d = 0; d < dout; ++d)
    for (int n = 0; n < w; ++n) blob[o_wt3 + (size_t)n * dout + d] = W3[(size_t)d * w + n];
  for (int r = 0; r < n_tidx; ++r)
    for (int n = 0; n < w; ++n) {
      double acc = h_b[0][n];
      if (has_emb)
        for (int e = 0; e < emb_dim; ++e)
          acc += (double)W0[(size_t)n * din_total + din + e] * (double)h_emb[(size_t)r * emb_dim + e];
      blob[o_b0 + (size_t)r * w + n] = (float)acc;
    }
  memcpy(&blob[o_b1], h_b[1], sizeof(float) * w);
  memcpy(&blob[o_b2], h_b[2], sizeof(float) * w);
  memcpy(&blob[o_b3], h_b[3], sizeof(float) * dout);
  for (int n = 0; n < w; ++n) {
    double u = 0.0, z = 0.0;
    for (int d = 0; d < dout; ++d) u += W3[(size_t)d * w + n];
    for (int d = 0; d < din - 1; ++d) z += W0[(size_t)n * din_total + d];
    blob[o_u3 + n] = (float)u;
    blob[o_zd1 + n] = (float)z;
  }

  if (fused_layout) {
    auto perm = [](int p) { const int h = p / 128, l = (p % 128) / 4, jj = p % 4; return l + 32 * (4 * h + jj); };
    for (int r = 0; r < w; ++r)
      for (int p = 0; p < w; ++p) {
        const int q = perm(p);
        blob[o_fwt1 + (size_t)r * w + p] = W1[(size_t)q * w + r];
        blob[o_fwt2 + (size_t)r * w + p] = W2[(size_t)q * w + r];
        blob[o_fw2r + (size_t)r * w + p] = W2[(size_t)r * w + q];
        blob[o_fw1r + (size_t)r * w + p] = W1[(size_t)r * w + q];
      }
  }
  if (tc_layout) {
    auto rn_tf32 = [](float x) {
      uint32_t u;
      memcpy(&u, &x, 4);
      u += 0x1000u;
      u &= 0xFFFFE000u;
      memcpy(&x, &u, 4);
      return x;
    };
    for (int g = 0; g < 4; ++g)
      for (int mt = 0; mt < tc_mt; ++mt)
        for (int kb = 0; kb < tc_kb; ++kb) {
          float* hi = &blob[o_tc + ((((size_t)g * tc_mt + mt) * tc_kb + kb) * 2 + 0) * 4096];
          float* lo = hi + 4096;
          for (int r = 0; r < 128; ++r)
            for (int kk = 0; kk < 32; ++kk) {
              const int row = mt * 128 + r, col = kb * 32 + kk;   // D[row][.] = sum_col A[row][col] * act[.][col]
              const float v = g == 0 ? W1[(size_t)row * w + col] : g == 1 ? W2[(size_t)row * w + col]
                            : g == 2 ? W2[(size_t)col * w + row] : W1[(size_t)col * w + row];
              const float h = rn_tf32(v);
              const size_t o = (size_t)r * 32 + ((((kk >> 2) ^ (r & 7)) << 2) | (kk & 3));
              hi[o] = h;
              lo[o] = rn_tf32(v - h);
            }
        }
  }
  if (wide_layout) {
    for (int g = 0; g < 4; ++g)
      for (int mt = 0; mt < tw_mt; ++mt)
        for (int kb = 0; kb < tw_kb; ++kb) {
          uint16_t* hi = reinterpret_cast<uint16_t*>(&blob[o_tw + ((((size_t)g * tw_mt + mt) * tw_kb + kb) * 2 + 0) * 4096]);
          uint16_t* lo = hi + 8192;
          for (int r = 0; r < 128; ++r)
            for (int kk = 0; kk < 64; ++kk) {
              const int row = mt * 128 + r, col = kb * 64 + kk;   // D[row][.] = sum_col A[row][col] * act[.][col]
              const float v = g == 0 ? W1[(size_t)row * w + col] : g == 1 ? W2[(size_t)row * w + col]
                            : g == 2 ? W2[(size_t)col * w + row] : W1[(size_t)col * w + row];
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn((v - __half2float(h)) * 2048.0f);
              const size_t o = (size_t)r * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
              memcpy(&hi[o], &h, 2);
              memcpy(&lo[o], &l, 2);
            }
        }
  }
  if (tc16_layout) {
    for (int g = 0; g < 4; ++g)
      for (int mt = 0; mt < 2; ++mt)
        for (int kb = 0; kb < 4; ++kb) {
          uint16_t* hi = reinterpret_cast<uint16_t*>(&blob[o_tc16 + ((((size_t)g * 2 + mt) * 4 + kb) * 2 + 0) * 4096]);
          uint16_t* lo = hi + 8192;
          for (int r = 0; r < 128; ++r)
            for (int kk = 0; kk < 64; ++kk) {
              // tile row r = TMEM lane: lanes {g, g+8, g+16, g+24} of a 32-lane quarter hold units 4s..4s+3, s = slot(g)
              // (the four lanes one thread reads with tcgen05.ld.16x128b), so a thread writes 4 consecutive k of the
              // next operand; slot() is the bank-conflict-free order of idff_sampler_tc16.cu
              const int gq = r & 7, slot = (gq & 1) | (((gq >> 1) & 1) << 2) | ((gq >> 2) << 1);
              const int unit = (r & ~31) + 4 * slot + ((r >> 3) & 3);
              const int row = mt * 128 + unit, col = kb * 64 + kk;   // D[row][.] = sum_col A[row][col] * act[.][col]
              const float v = g == 0 ? W1[(size_t)row * w + col] : g == 1 ? W2[(size_t)row * w + col]
                            : g == 2 ? W2[(size_t)col * w + row] : W1[(size_t)col * w + row];
              const __half h = __float2half_rn(v);
              const __half l = __float2half_rn((v - __half2float(h)) * 2048.0f);
              const size_t o = (size_t)r * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
              memcpy(&hi[o], &h, 2);
              memcpy(&lo[o], &l, 2);
            }
        }
  }
  for (int n = 0; n < w; ++n)
    for (int d = 0; d < din - 1; ++d) blob[o_w0n + (size_t)n * (din - 1) + d] = W0[(size_t)n * din_total + d];

  if (md16_layout) {
    const int Dm = din - 1;
    for (int t = 0; t < 13; ++t) {
      uint16_t* hi = reinterpret_cast<uint16_t*>(&blob[o_md16 + (size_t)t * 2 * 4096]);
      uint16_t* lo = hi + 8192;
      const int kb = t == 0 ? 0 : (t - 1) & 1;   // k-block of 64 inside the GEMM
      const int gm = t == 0 ? 0 : 1 + (t - 1) / 2;   // 0: W0, 1: W1, 2: W2, 3: W3, 4: W2^T, 5: W1^T, 6: W0^T
      for (int r = 0; r < 128; ++r)
        for (int kk = 0; kk < 64; ++kk) {
          const int gq = r & 7, slot = (gq & 1) | (((gq >> 1) & 1) << 2) | ((gq >> 2) << 1);
          const int row = (r & ~31) + 4 * slot + ((r >> 3) & 3);   // TMEM lane r holds output row `row` (see tc16_tiles)
          const int col = kb * 64 + kk;
          float v = 0.0f;
          switch (gm) {
            case 0: v = col < Dm ? W0[(size_t)row * din_total + col] : 0.0f; break;
            case 1: v = W1[(size_t)row * w + col]; break;
            case 2: v = W2[(size_t)row * w + col]; break;
            case 3: v = row < dout ? W3[(size_t)row * w + col] : 0.0f; break;
            case 4: v = W2[(size_t)col * w + row]; break;
            case 5: v = W1[(size_t)col * w + row]; break;
            default: v = row < Dm ? W0[(size_t)col * din_total + row] : 0.0f; break;
          }
          const __half h = __float2half_rn(v);
          const __half l = __float2half_rn((v - __half2float(h)) * 2048.0f);
          const size_t o = (size_t)r * 64 + ((((kk >> 3) ^ (r & 7)) << 3) | (kk & 7));
          memcpy(&hi[o], &h, 2);
          memcpy(&lo[o], &l, 2);
        }
    }
  }

  int prev = 0;
  IDFF_CUDA(cudaGetDevice(&prev));
  IDFF_CUDA(cudaSetDevice(device));
  float* d_blob = nullptr;
  cudaError_t e = cudaMalloc(&d_blob, off * sizeof(float));
  if (e != cudaSuccess) {
    cudaSetDevice(prev);
    set_error("idff_weights_create: cudaMalloc(%zu) failed: %s", off * sizeof(float), cudaGetErrorString(e));
    return IDFF_ENOMEM;
  }
  e = cudaMemcpy(d_blob, blob.data(), off * sizeof(float), cudaMemcpyHostToDevice);
  unsigned long long* d_counters = nullptr;
  if (e == cudaSuccess) e = cudaMalloc(&d_counters, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(d_counters, 0, 2 * sizeof(unsigned long long));
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    cudaFree(d_blob);
    if (d_counters) cudaFree(d_counters);
    set_error("idff_weights_create: cudaMemcpy failed: %s", cudaGetErrorString(e));
    return IDFF_ECUDA;
  }
  idff_weights* h = new idff_weights();
  h->device = device;
  for (int i = 0; i < 5; ++i) h->dims[i] = dims[i];
  h->emb_rows = has_emb ? emb_rows : 0;
  h->emb_dim = has_emb ? emb_dim : 0;
  h->d_blob = d_blob;
  h->blob_floats = off;
  NetView& v = h->v;
  v.din = din; v.din_pad = din_pad; v.w = w; v.dout = dout; v.n_tidx = n_tidx;
  v.wt0 = d_blob + o_wt0; v.wt1 = d_blob + o_wt1; v.wt2 = d_blob + o_wt2;
  v.w1 = d_blob + o_w1; v.w2 = d_blob + o_w2; v.w3 = d_blob + o_w3; v.wt3 = d_blob + o_wt3;
  v.b0 = d_blob + o_b0; v.b1 = d_blob + o_b1; v.b2 = d_blob + o_b2; v.b3 = d_blob + o_b3;
  v.u3 = d_blob + o_u3; v.zd1 = d_blob + o_zd1;
  v.fwt1 = fused_layout ? d_blob + o_fwt1 : nullptr; v.fwt2 = fused_layout ? d_blob + o_fwt2 : nullptr;
  v.fw2r = fused_layout ? d_blob + o_fw2r : nullptr; v.fw1r = fused_layout ? d_blob + o_fw1r : nullptr;
  v.w0n = d_blob + o_w0n;
  v.tc_tiles = tc_layout ? d_blob + o_tc : nullptr;
  v.tc16_tiles = tc16_layout ? d_blob + o_tc16 : nullptr;
  v.tcw16_tiles = wide_layout ? d_blob + o_tw : nullptr;
  v.md16_tiles = md16_layout ? d_blob + o_md16 : nullptr;
  h->d_stash = nullptr; h->stash_bytes = 0; h->tc_verified = 0;
  h->d_sweep = nullptr; h->sweep_bytes = 0;
  h->d_sweep_out = nullptr; h->sweep_out_bytes = 0;
  h->d_fix_x = nullptr; h->fix_x_bytes = 0; h->d_fix_tab = nullptr; h->fix_tab_bytes = 0;
  h->d_counters = d_counters;
  *out = h;
  return IDFF_OK;
}

extern "C" int idff_weights_destroy(idff_weights_t* w) {
  if (!w) return IDFF_OK;
  cudaFree(w->d_blob);
  if (w->d_stash) cudaFree(w->d_stash);
  if (w->d_sweep) cudaFree(w->d_sweep);
  if (w->d_sweep_out) cudaFree(w->d_sweep_out);
  if (w->d_counters) cudaFree(w->d_counters);
  if (w->d_fix_x) cudaFree(w->d_fix_x);
  if (w->d_fix_tab) cudaFree(w->d_fix_tab);
  delete w;
  return IDFF_OK;
}
