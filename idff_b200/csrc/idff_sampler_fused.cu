// Persistent warp-tiled IDFF sampler for sm_100a (IDFF_IMPL_FUSED): hidden width 128 or 256.
//
// CTA = 8 consumer warps + 1 producer warp, one CTA per SM, looping over tiles of 32 particles.
//   * A consumer warp owns 4 particles for the whole solve.  Its activations (all jets of the 4 particles, all w
//     units of the current layer) live in a warp-private shared-memory panel A[k][row], row = jet*4 + particle;
//     lane l owns the units {l, l+32, ...} of every layer, so the per-unit SELU derivative data a reverse sweep
//     needs never leaves the thread that produced it (layer 3: registers; layer 2: a per-thread stash; layer 1:
//     recomputed from x).  Warps never synchronise with each other.
//   * Every w x w layer is a register-tiled FMA GEMM  acc[row][unit] += A[k][row] * B[k][unit]  (8..12 rows x w/32
//     units per thread).  B is W^T for the forward passes and W for the reverse sweep, streamed from L2 in 16-row
//     chunks by the producer warp with TMA bulk copies (cp.async.bulk + mbarrier complete_tx) through a 3/4-stage
//     ring shared by the 8 consumer warps (full/empty mbarriers; no __syncthreads in the steady state).
//   * The RK4 (3/8 rule) / SRK stage updates, the t == 0 / t == 1 branches and the gamma combination happen in the
//     warp between GEMMs; particle state stays in shared memory across all stages of the solve.
// fp32 FMA throughout: the 1e-5 per-evaluation contract rules out single-pass TF32/BF16 tensor-core MMA.
#include "idff_common.cuh"

namespace idff {
void jets_needed(const idff_field_params_t& p, int* nj, int* reverse);

namespace fz {

constexpr int KC = 16;                    // weight rows per pipeline chunk
constexpr int NCW = 8;                    // consumer warps
constexpr int PT = 4;                     // particles per consumer warp
constexpr int TILE = NCW * PT;            // particles per CTA iteration
constexpr int NTHREADS = (NCW + 1) * 32;  // + producer warp
constexpr int NARR = 9;                   // per-warp state arrays

template <int W>
struct Geo {
  static constexpr int CT = W / 32;                 // units per thread
  static constexpr int NSTAGE = W == 256 ? 3 : 4;   // ring depth
  static constexpr int CHUNK = KC * W;              // floats per chunk
  static constexpr int NCHUNK = W / KC;             // chunks per GEMM
};

// ---- PTX helpers (mbarrier + TMA bulk copy) -------------------------------------------------------------------------
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(s32(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(s32(b)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}

// d += a * b on both halves of a 64-bit register pair (SASS FFMA2)
__device__ __forceinline__ void ffma2(float2& d, float2 a, float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  const unsigned long long aa = *reinterpret_cast<unsigned long long*>(&a), bb = *reinterpret_cast<unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  d = *reinterpret_cast<float2*>(&dd);
}

// 16-byte slot (= one jet of the warp's 4 particles) of unit k inside the warp panel.  With 2 jets the two slots of
// a unit are swapped on every other group of 4 units so the epilogue's unit-stride 128-bit stores hit distinct banks.
template <int NJ>
__device__ __forceinline__ int slot_of(int k, int s) {
  return NJ == 2 ? (s ^ ((k >> 2) & 1)) : s;
}

struct FParams {
  NetView net;
  int variant, D, tidx, reverse, mode;   // mode 0 eval, 1 rk4, 2 sde
  long N;
  const float* x_in;     // [N][D] initial state (rk4/sde) or evaluation points (eval)
  const float* x0_cfm;   // eval mode, CFM variant, t != 0
  float* out;            // rk4: final [N][D]; eval: f [N][D]; sde: [n_ts][N][D]
  float* traj;           // rk4: optional [n_grid][N][D] (row 0 written by the host wrapper)
  const float* Ik;
  const float* Ik0;
  float4* stash_g;       // global per-thread stash (order 3) or null -> shared memory
};

struct Layout {   // offsets in floats from the dynamic smem base
  int bars, stages, A, stash, wt0, vec, w3, b3, w0n, state, sd, total;
};
__host__ __device__ inline Layout make_layout(int W, int NJ, int D, int din_pad, int dout, bool dsmall, bool stash_smem) {
  const int CT = W / 32, NSTAGE = W == 256 ? 3 : 4;
  Layout L;
  int o = 0;
  L.bars = o; o += 32;
  L.stages = o; o += NSTAGE * KC * W;
  L.A = o; o += NCW * W * 4 * NJ;
  L.stash = o; o += stash_smem ? NCW * 32 * NJ * PT * CT : 0;
  L.wt0 = o; o += din_pad * W;
  L.vec = o; o += 5 * W;                       // b0, b1, b2, u3, zd1
  L.w3 = o; o += dout * W;
  L.b3 = o; o += (dout + 3) & ~3;
  L.w0n = o; o += dsmall ? 0 : W * D;
  L.sd = (PT * (D > dout ? D : dout) + 3) & ~3;
  L.state = o; o += NCW * NARR * L.sd;
  L.total = o;
  return L;
}

__device__ __forceinline__ void adjoint3(int nja, float code, float zd, float zdd, float A0, float A1, float A2,
                                         float& P, float& Q, float& R) {
  float s1, E;
  selu_decode(code, s1, E);
  P = A0 * s1;
  Q = 0.0f;
  R = 0.0f;
  if (nja >= 2) {
    P = fmaf(A1 * E, zd, P);
    Q = A1 * s1;
  }
  if (nja >= 3) {
    P = fmaf(A2 * E, fmaf(zd, zd, zdd), P);
    Q = fmaf(2.0f * A2 * E, zd, Q);
    R = A2 * s1;
  }
}

// ---- one consumer warp ---------------------------------------------------------------------------------------------
template <int W, int NJ, int DT>
struct Warp {
  static constexpr int CT = Geo<W>::CT, RW = 4 * NJ, NSTAGE = Geo<W>::NSTAGE;
  const float* stages;
  uint64_t *full, *empty;
  float* A;
  float4* stash;   // indexed [slot * 256 + ctid]
  const float *wt0, *b0, *b1, *b2, *u3, *zd1, *w3, *b3, *w0n;
  float *y, *k1, *k2, *k3, *xs, *fb, *x0c, *pred, *g;
  int lane, ctid, D, dout, variant, reverse;
  int stage;
  uint32_t phase;
  bool ready;   // the current chunk's full barrier was already observed complete

  __device__ __forceinline__ int dim() const { return DT > 0 ? DT : D; }

  // acc[row][j] = sum_k A[k][row] * B[k][unit_j]; consumes W/KC chunks of the shared weight ring.
  // The inner product runs on packed FFMA2 (fma.rn.f32x2: two IEEE fp32 FMAs per instruction on a 64-bit register
  // pair, the row operand broadcast): measured 105 vs 90 FMA/clk/SM for the scalar form (profiles/README.md).
  template <int NJA>
  __device__ __forceinline__ void gemm(float (&acc)[NJA * 4][CT]) {
    float2 acc2[NJA * 4][CT / 2];
#pragma unroll
    for (int r = 0; r < NJA * 4; ++r)
#pragma unroll
      for (int j = 0; j < CT / 2; ++j) acc2[r][j] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int ch = 0; ch < Geo<W>::NCHUNK; ++ch) {
      if (!ready) mbar_wait(&full[stage], phase);
      const float* Bs = stages + stage * Geo<W>::CHUNK + lane * 4;
      const float* As = A + ch * KC * RW;
      const int nstage = stage + 1 == NSTAGE ? 0 : stage + 1;
      const uint32_t nphase = stage + 1 == NSTAGE ? phase ^ 1u : phase;
#pragma unroll
      for (int kk = 0; kk < KC; ++kk) {
        // probe the next chunk's barrier in the middle of this one so its latency hides behind the FMAs
        if (kk == KC / 2) ready = mbar_test_wait(&full[nstage], nphase);
        float2 b2[CT / 2];
        float a[NJA * 4];
#pragma unroll
        for (int h = 0; h < CT / 4; ++h) {
          const float4 v = *reinterpret_cast<const float4*>(Bs + kk * W + h * 128);
          b2[2 * h + 0] = make_float2(v.x, v.y);
          b2[2 * h + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int s = 0; s < NJA; ++s) {
          const float4 v = *reinterpret_cast<const float4*>(As + kk * RW + slot_of<NJ>(kk, s) * 4);
          a[4 * s + 0] = v.x; a[4 * s + 1] = v.y; a[4 * s + 2] = v.z; a[4 * s + 3] = v.w;
        }
#pragma unroll
        for (int r = 0; r < NJA * 4; ++r)
#pragma unroll
          for (int j = 0; j < CT / 2; ++j) ffma2(acc2[r][j], make_float2(a[r], a[r]), b2[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
      if (++stage == NSTAGE) { stage = 0; phase ^= 1u; }
    }
#pragma unroll
    for (int r = 0; r < NJA * 4; ++r)
#pragma unroll
      for (int j = 0; j < CT / 2; ++j) {
        acc[r][2 * j] = acc2[r][j].x;
        acc[r][2 * j + 1] = acc2[r][j].y;
      }
  }

  __device__ __forceinline__ void put(int n, int s, float v0, float v1, float v2, float v3) {
    *reinterpret_cast<float4*>(A + n * RW + slot_of<NJ>(n, s) * 4) = make_float4(v0, v1, v2, v3);
  }

  // layer-1 pre-activations of the warp's 4 particles at this lane's units (K = D+1: computed directly)
  __device__ __forceinline__ void l1_preact(float t, float (&z)[4][CT]) {
    const int Dd = dim();
#pragma unroll
    for (int j = 0; j < CT; ++j) {
      const int n = lane + 32 * j;
      const float zb = fmaf(wt0[Dd * W + n], t, b0[n]);
#pragma unroll
      for (int p = 0; p < 4; ++p) z[p][j] = zb;
    }
    if (DT > 0) {
#pragma unroll
      for (int d = 0; d < (DT > 0 ? DT : 1); ++d) {
        float xv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) xv[p] = xs[p * DT + d];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const float wv = wt0[d * W + lane + 32 * j];
#pragma unroll
          for (int p = 0; p < 4; ++p) z[p][j] = fmaf(wv, xv[p], z[p][j]);
        }
      }
    } else {
#pragma unroll 2
      for (int d = 0; d < Dd; ++d) {
        float xv[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) xv[p] = xs[p * Dd + d];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const float wv = wt0[d * W + lane + 32 * j];
#pragma unroll
          for (int p = 0; p < 4; ++p) z[p][j] = fmaf(wv, xv[p], z[p][j]);
        }
      }
    }
  }

  // out[p][d] = bias[d] + sum_k A[k][slot 0][p] * M[k*nd + d]   (wide-D path: lanes over d)
  __device__ __forceinline__ void small_from_panel(const float* M, int nd, const float* bias, float* out) {
    for (int d = lane; d < nd; d += 32) {
      float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll 4
      for (int k = 0; k < W; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(A + k * RW + slot_of<NJ>(k, 0) * 4);
        const float m = M[k * nd + d];
        s0 = fmaf(a.x, m, s0); s1 = fmaf(a.y, m, s1); s2 = fmaf(a.z, m, s2); s3 = fmaf(a.w, m, s3);
      }
      const float bb = bias ? bias[d] : 0.0f;
      out[0 * nd + d] = s0 + bb; out[1 * nd + d] = s1 + bb; out[2 * nd + d] = s2 + bb; out[3 * nd + d] = s3 + bb;
    }
  }

  // register partials part[p][d] (this lane's units) -> totals in shared memory (narrow-D path)
  __device__ __forceinline__ void reduce_store(float (&part)[4][DT > 0 ? DT : 1], const float* bias, float* out) {
    constexpr int DD = DT > 0 ? DT : 1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int d = 0; d < DD; ++d) part[p][d] += __shfl_xor_sync(0xffffffffu, part[p][d], o);
    float v = 0.0f;
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int d = 0; d < DD; ++d)
        if (lane == p * DD + d) v = part[p][d] + (bias ? bias[d] : 0.0f);
    if (lane < 4 * DD) out[lane] = v;
  }

  // One evaluation of the field at the warp's xs -> fb.  NJA jets; HEAVY = with the reverse sweep.
  template <int NJA, bool HEAVY>
  __device__ __forceinline__ void eval(const Stage& st) {
    const int Dd = dim();
    float acc[NJA * 4][CT];
    // ---- layer 1
    {
      float z[4][CT];
      l1_preact(st.t, z);
#pragma unroll
      for (int j = 0; j < CT; ++j) {
        const int n = lane + 32 * j;
        float a[4], ad[4], add[4];
        const float zd = NJA >= 2 ? zd1[n] : 0.0f;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          float code, s1, E;
          a[p] = selu_fwd(z[p][j], code);
          selu_decode(code, s1, E);
          ad[p] = s1 * zd;
          add[p] = E * zd * zd;
        }
        put(n, 0, a[0], a[1], a[2], a[3]);
        if (NJA >= 2) put(n, 1, ad[0], ad[1], ad[2], ad[3]);
        if (NJA >= 3) put(n, 2, add[0], add[1], add[2], add[3]);
      }
    }
    __syncwarp();
    float part[4][DT > 0 ? DT : 1];
    constexpr int NG = HEAVY ? 4 : 2;
#pragma unroll 1
    for (int gi = 0; gi < NG; ++gi) {
      gemm<NJA>(acc);
      if (gi == 0) {
        // ---- layer 2 forward epilogue: activations of all jets; stash (code, z', z'') for the reverse sweep
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int n = lane + 32 * j;
          const float bb = b1[n];
          float a[4], ad[4], add[4], cd[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float s1, E;
            a[p] = selu_fwd(acc[p][j] + bb, cd[p]);
            selu_decode(cd[p], s1, E);
            ad[p] = NJA >= 2 ? s1 * acc[(NJA >= 2 ? 4 : 0) + p][j] : 0.0f;
            add[p] = NJA >= 3 ? fmaf(E * acc[(NJA >= 2 ? 4 : 0) + p][j], acc[(NJA >= 2 ? 4 : 0) + p][j],
                                     s1 * acc[(NJA >= 3 ? 8 : 0) + p][j])
                              : 0.0f;
          }
          if (HEAVY) {
            stash[(j * NJ + 0) * 256 + ctid] = make_float4(cd[0], cd[1], cd[2], cd[3]);
            if (NJA >= 2)
              stash[(j * NJ + 1) * 256 + ctid] = make_float4(acc[(NJA >= 2 ? 4 : 0) + 0][j], acc[(NJA >= 2 ? 4 : 0) + 1][j],
                                                             acc[(NJA >= 2 ? 4 : 0) + 2][j], acc[(NJA >= 2 ? 4 : 0) + 3][j]);
            if (NJA >= 3)
              stash[(j * NJ + 2) * 256 + ctid] = make_float4(acc[(NJA >= 3 ? 8 : 0) + 0][j], acc[(NJA >= 3 ? 8 : 0) + 1][j],
                                                             acc[(NJA >= 3 ? 8 : 0) + 2][j], acc[(NJA >= 3 ? 8 : 0) + 3][j]);
          }
          put(n, 0, a[0], a[1], a[2], a[3]);
          if (NJA >= 2) put(n, 1, ad[0], ad[1], ad[2], ad[3]);
          if (NJA >= 3) put(n, 2, add[0], add[1], add[2], add[3]);
        }
      } else if (gi == 1) {
        // ---- layer 3 forward epilogue: x1hat = W3 a3 + b3; adjoint seeds of c1*phi + c2*D_d phi + c3*D_d^2 phi
        if (DT > 0) {
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : 1); ++d) part[p][d] = 0.0f;
        }
        float a3[4][CT];
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int n = lane + 32 * j;
          const float bb = b2[n];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float cd;
            a3[p][j] = selu_fwd(acc[p][j] + bb, cd);
            acc[p][j] = cd;   // keep the derivative code in place of z
          }
          if (DT > 0) {
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : 1); ++d) {
              const float wv = w3[d * W + n];
#pragma unroll
              for (int p = 0; p < 4; ++p) part[p][d] = fmaf(a3[p][j], wv, part[p][d]);
            }
          } else {
            put(n, 0, a3[0][j], a3[1][j], a3[2][j], a3[3][j]);
          }
        }
        if (DT > 0) {
          reduce_store(part, b3, pred);
        } else {
          __syncwarp();
          small_from_panel(w3, dout, b3, pred);   // w3 holds W3^T [W][dout] on this path
          __syncwarp();
        }
        if (HEAVY) {
#pragma unroll
          for (int j = 0; j < CT; ++j) {
            const int n = lane + 32 * j;
            const float u = u3[n];
            const float A0 = st.c[0] * u, A1 = st.c[1] * u, A2 = st.c[2] * u;
            float P[4], Q[4], R[4];
#pragma unroll
            for (int p = 0; p < 4; ++p)
              adjoint3(NJA, acc[p][j], NJA >= 2 ? acc[(NJA >= 2 ? 4 : 0) + p][j] : 0.0f,
                       NJA >= 3 ? acc[(NJA >= 3 ? 8 : 0) + p][j] : 0.0f, A0, A1, A2, P[p], Q[p], R[p]);
            put(n, 0, P[0], P[1], P[2], P[3]);
            if (NJA >= 2) put(n, 1, Q[0], Q[1], Q[2], Q[3]);
            if (NJA >= 3) put(n, 2, R[0], R[1], R[2], R[3]);
          }
        }
      } else if (gi == 2) {
        // ---- reverse, layer 3 -> 2: acc = adjoints of (a2, a2', a2''); combine with the stash
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int n = lane + 32 * j;
          const float4 cd = stash[(j * NJ + 0) * 256 + ctid];
          const float4 zd = NJA >= 2 ? stash[(j * NJ + (NJA >= 2 ? 1 : 0)) * 256 + ctid] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 zdd = NJA >= 3 ? stash[(j * NJ + (NJA >= 3 ? 2 : 0)) * 256 + ctid] : make_float4(0.f, 0.f, 0.f, 0.f);
          const float cda[4] = {cd.x, cd.y, cd.z, cd.w}, zda[4] = {zd.x, zd.y, zd.z, zd.w},
                      zdda[4] = {zdd.x, zdd.y, zdd.z, zdd.w};
          float P[4], Q[4], R[4];
#pragma unroll
          for (int p = 0; p < 4; ++p)
            adjoint3(NJA, cda[p], zda[p], zdda[p], acc[p][j], NJA >= 2 ? acc[(NJA >= 2 ? 4 : 0) + p][j] : 0.0f,
                     NJA >= 3 ? acc[(NJA >= 3 ? 8 : 0) + p][j] : 0.0f, P[p], Q[p], R[p]);
          put(n, 0, P[0], P[1], P[2], P[3]);
          if (NJA >= 2) put(n, 1, Q[0], Q[1], Q[2], Q[3]);
          if (NJA >= 3) put(n, 2, R[0], R[1], R[2], R[3]);
        }
      } else {
        // ---- reverse, layer 2 -> 1 -> x: layer-1 derivative data recomputed from x; g = W0x^T P1
        float z[4][CT];
        l1_preact(st.t, z);
        if (DT > 0) {
#pragma unroll
          for (int p = 0; p < 4; ++p)
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : 1); ++d) part[p][d] = 0.0f;
        }
#pragma unroll
        for (int j = 0; j < CT; ++j) {
          const int n = lane + 32 * j;
          const float zd = zd1[n];
          float P[4];
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float cd, Q, R;
            (void)selu_fwd(z[p][j], cd);
            adjoint3(NJA, cd, zd, 0.0f, acc[p][j], NJA >= 2 ? acc[(NJA >= 2 ? 4 : 0) + p][j] : 0.0f,
                     NJA >= 3 ? acc[(NJA >= 3 ? 8 : 0) + p][j] : 0.0f, P[p], Q, R);
          }
          if (DT > 0) {
#pragma unroll
            for (int d = 0; d < (DT > 0 ? DT : 1); ++d) {
              const float wv = wt0[d * W + n];
#pragma unroll
              for (int p = 0; p < 4; ++p) part[p][d] = fmaf(P[p], wv, part[p][d]);
            }
          } else {
            put(n, 0, P[0], P[1], P[2], P[3]);
          }
        }
        if (DT > 0) {
          reduce_store(part, nullptr, g);
        } else {
          __syncwarp();
          small_from_panel(w0n, Dd, nullptr, g);
        }
      }
      __syncwarp();
    }
    // ---- combine (reference operation order: (a*(x1hat-x))/(1-t) + momentum)
    for (int o = lane; o < 4 * Dd; o += 32) {
      const float x = xs[o], pr = pred[o];
      float f;
      if (variant == IDFF_FIELD_CFM) {
        if (st.kind == 0) x0c[o] = x;
        f = __fsub_rn(pr, x0c[o]);
      } else {
        const float diff = __fsub_rn(pr, x);
        if (st.kind == 0) f = diff;
        else if (st.kind == 1) f = __fdiv_rn(diff, st.end_div);
        else {
          f = __fdiv_rn(__fmul_rn(st.a, diff), st.den);
          if (HEAVY) f = __fadd_rn(f, g[o]);
        }
      }
      fb[o] = f;
    }
    __syncwarp();
  }

  __device__ __forceinline__ void eval_any(const Stage& st) {
    if (st.kind == 2 && reverse) eval<NJ, true>(st);
    else eval<1, false>(st);
  }
};

template <int W, int NJ, int DT>
__global__ void __launch_bounds__(NTHREADS, 1)
k_fused(const __grid_constant__ FParams P, const __grid_constant__ Rk4Table rk, const __grid_constant__ SdeTable sde,
        const __grid_constant__ Stage one) {
  extern __shared__ __align__(128) float smem[];
  using G = Geo<W>;
  const int D = P.D, dout = P.net.dout;
  const bool stash_smem = P.stash_g == nullptr;
  const Layout L = make_layout(W, NJ, D, P.net.din_pad, dout, DT > 0, stash_smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.bars);
  uint64_t* empty = full + G::NSTAGE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- one-time setup: barriers, resident small operands
  if (tid == 0) {
    for (int s = 0; s < G::NSTAGE; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW);
    }
    fence_mbar_init();
  }
  for (int i = tid; i < P.net.din_pad * W; i += NTHREADS) smem[L.wt0 + i] = P.net.wt0[i];
  for (int i = tid; i < W; i += NTHREADS) {
    smem[L.vec + 0 * W + i] = P.net.b0[(size_t)P.tidx * W + i];
    smem[L.vec + 1 * W + i] = P.net.b1[i];
    smem[L.vec + 2 * W + i] = P.net.b2[i];
    smem[L.vec + 3 * W + i] = P.net.u3[i];
    smem[L.vec + 4 * W + i] = P.net.zd1[i];
  }
  {
    const float* w3src = DT > 0 ? P.net.w3 : P.net.wt3;   // [dout][W] (narrow D) or [W][dout] (wide D)
    for (int i = tid; i < dout * W; i += NTHREADS) smem[L.w3 + i] = w3src[i];
    for (int i = tid; i < dout; i += NTHREADS) smem[L.b3 + i] = P.net.b3[i];
    if (DT == 0)
      for (int i = tid; i < W * D; i += NTHREADS) smem[L.w0n + i] = P.net.w0n[i];
  }
  __syncthreads();

  // evaluation schedule of one tile (identical for every warp of the CTA, producer included)
  const int n_eval = P.mode == 0 ? 1 : (P.mode == 1 ? 4 * rk.n_iv : sde.n_steps * (sde.method == 1 ? 1 : 3));
  auto stage_of = [&](int e) -> const Stage& {
    if (P.mode == 0) return one;
    if (P.mode == 1) return rk.st[e];
    const int per = sde.method == 1 ? 1 : 3;
    return sde.st[e / per][e % per];
  };

  if (warp == NCW) {
    // ===== producer warp: stream the weight chunks of every GEMM of every evaluation of every tile
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long tile = blockIdx.x; tile * TILE < P.N; tile += gridDim.x) {
        for (int e = 0; e < n_eval; ++e) {
          const Stage& st = stage_of(e);
          const int ng = (st.kind == 2 && P.reverse) ? 4 : 2;
          for (int gi = 0; gi < ng; ++gi) {
            const float* src = gi == 0 ? P.net.fwt1 : (gi == 1 ? P.net.fwt2 : (gi == 2 ? P.net.fw2r : P.net.fw1r));
            for (int ch = 0; ch < G::NCHUNK; ++ch) {
              while (!mbar_try_wait(&empty[stage], phase ^ 1u)) __nanosleep(128);   // do not steal issue slots
              mbar_arrive_expect_tx(&full[stage], G::CHUNK * 4);
              tma_load_1d(smem + L.stages + stage * G::CHUNK, src + (size_t)ch * G::CHUNK, G::CHUNK * 4, &full[stage]);
              if (++stage == G::NSTAGE) { stage = 0; phase ^= 1u; }
            }
          }
        }
      }
    }
    return;
  }

  // ===== consumer warps
  Warp<W, NJ, DT> wp;
  wp.stages = smem + L.stages;
  wp.full = full;
  wp.empty = empty;
  wp.A = smem + L.A + warp * (W * 4 * NJ);
  wp.stash = stash_smem ? reinterpret_cast<float4*>(smem + L.stash)
                        : P.stash_g + (size_t)blockIdx.x * (NJ * G::CT * 256);
  wp.wt0 = smem + L.wt0;
  wp.b0 = smem + L.vec; wp.b1 = wp.b0 + W; wp.b2 = wp.b1 + W; wp.u3 = wp.b2 + W; wp.zd1 = wp.u3 + W;
  wp.w3 = smem + L.w3; wp.b3 = smem + L.b3; wp.w0n = smem + L.w0n;
  float* stt = smem + L.state + warp * (NARR * L.sd);
  wp.y = stt; wp.k1 = stt + L.sd; wp.k2 = stt + 2 * L.sd; wp.k3 = stt + 3 * L.sd; wp.xs = stt + 4 * L.sd;
  wp.fb = stt + 5 * L.sd; wp.x0c = stt + 6 * L.sd; wp.pred = stt + 7 * L.sd; wp.g = stt + 8 * L.sd;
  wp.lane = lane; wp.ctid = tid; wp.D = D; wp.dout = dout; wp.variant = P.variant; wp.reverse = P.reverse;
  wp.stage = 0; wp.phase = 0; wp.ready = false;
  const int Dd = wp.dim();
  const float third = (float)(1.0 / 3.0);

  for (long tile = blockIdx.x; tile * TILE < P.N; tile += gridDim.x) {
    const long base = tile * TILE + warp * PT;
    for (int o = lane; o < 4 * Dd; o += 32) {
      const long row = base + o / Dd;
      const float v = row < P.N ? P.x_in[row * Dd + (o % Dd)] : 0.0f;
      wp.y[o] = v;
      wp.xs[o] = v;
      if (P.variant == IDFF_FIELD_CFM && P.x0_cfm) wp.x0c[o] = row < P.N ? P.x0_cfm[row * Dd + (o % Dd)] : 0.0f;
      if (P.mode == 2 && row < P.N) P.out[row * Dd + (o % Dd)] = v;   // ys[0] = y0
    }
    __syncwarp();
    // one evaluation loop for all three modes (a single inlined copy of the field evaluation)
    float* yprev = wp.k3;   // sde: previous step's state for the interpolated outputs
    int emit = 0;
    while (P.mode == 2 && emit < sde.n_emit && sde.emit_step[emit] < 0) ++emit;
    const int per = sde.method == 1 ? 1 : 3;
#pragma unroll 1
    for (int e = 0; e < n_eval; ++e) {
      wp.eval_any(stage_of(e));
      if (P.mode == 0) {
        for (int o = lane; o < 4 * Dd; o += 32) {
          const long row = base + o / Dd;
          if (row < P.N) P.out[row * Dd + (o % Dd)] = wp.fb[o];
        }
      } else if (P.mode == 1) {
        // torchdiffeq rk4 == rk4_alt_step_func (3/8 rule), operation order kept
        const int iv = e >> 2, sg = e & 3;
        const float dt = rk.dt[iv];
        for (int o = lane; o < 4 * Dd; o += 32) {
          const float y = wp.y[o], f = wp.fb[o];
          float xn;
          if (sg == 0) {
            wp.k1[o] = f;
            xn = __fadd_rn(y, __fmul_rn(__fmul_rn(dt, f), third));
          } else if (sg == 1) {
            wp.k2[o] = f;
            xn = __fadd_rn(y, __fmul_rn(dt, __fsub_rn(f, __fmul_rn(wp.k1[o], third))));
          } else if (sg == 2) {
            wp.k3[o] = f;
            xn = __fadd_rn(y, __fmul_rn(dt, __fadd_rn(__fsub_rn(wp.k1[o], wp.k2[o]), f)));
          } else {
            const float sum = __fadd_rn(__fadd_rn(wp.k1[o], __fmul_rn(3.0f, __fadd_rn(wp.k2[o], wp.k3[o]))), f);
            xn = __fadd_rn(y, __fmul_rn(__fmul_rn(sum, dt), 0.125f));
            wp.y[o] = xn;
            const long row = base + o / Dd;
            if (P.traj != nullptr && row < P.N) P.traj[((size_t)(iv + 1) * P.N + row) * Dd + (o % Dd)] = xn;
            if (e == n_eval - 1 && row < P.N) P.out[row * Dd + (o % Dd)] = xn;
          }
          wp.xs[o] = xn;
        }
      } else {
        // torchsde srk (SRI2, diagonal, state-independent g) / euler with injected increments
        const int stp = e / per, sg = e - stp * per;
        const float h = sde.h[stp];
        const size_t noff = (size_t)stp * P.N * Dd;
        const float rdt = __fdiv_rn(1.0f, h), sq = __fsqrt_rn(h);
        for (int o = lane; o < 4 * Dd; o += 32) {
          const long row = base + o / Dd;
          const size_t gi = noff + (size_t)row * Dd + (o % Dd);
          const float y = wp.y[o], f = wp.fb[o];
          if (sde.method == 1) {
            const float ik = row < P.N ? P.Ik[gi] : 0.0f;
            const float yn = __fadd_rn(__fadd_rn(y, __fmul_rn(f, h)), __fmul_rn(sde.gval[stp][0], ik));
            yprev[o] = y; wp.y[o] = yn; wp.xs[o] = yn;
          } else if (sg == 0) {
            wp.k1[o] = f;
            const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
            const float t1 = __fmul_rn(__fmul_rn(1.0f, f), h);
            const float t2 = __fmul_rn(__fmul_rn(__fmul_rn(0.0f, sde.gval[stp][0]), ik0), rdt);
            wp.xs[o] = __fadd_rn(__fadd_rn(y, t1), t2);
          } else if (sg == 1) {
            wp.k2[o] = f;
            const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
            float v = y;
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, wp.k1[o]), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(1.0f, sde.gval[stp][0]), ik0), rdt));
            v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(0.25f, f), h)),
                          __fmul_rn(__fmul_rn(__fmul_rn(0.5f, sde.gval[stp][1]), ik0), rdt));
            wp.xs[o] = v;
          } else {
            const float ik = row < P.N ? P.Ik[gi] : 0.0f;
            const float ik0 = row < P.N ? P.Ik0[gi] : 0.0f;
            const float ikk = __fmul_rn(__fsub_rn(__fmul_rn(ik, ik), h), 0.5f);
            const float ikkk = __fdiv_rn(__fsub_rn(__fmul_rn(__fmul_rn(ik, ik), ik), __fmul_rn(__fmul_rn(3.0f, h), ik)), 6.0f);
            const float b1[4] = {-1.0f, (float)(4.0 / 3.0), (float)(2.0 / 3.0), 0.0f};
            const float b2[4] = {1.0f, (float)(-4.0 / 3.0), (float)(1.0 / 3.0), 0.0f};
            const float b3[4] = {2.0f, (float)(-4.0 / 3.0), (float)(-2.0 / 3.0), 0.0f};
            const float b4[4] = {-2.0f, (float)(5.0 / 3.0), (float)(-2.0 / 3.0), 1.0f};
            const float al[4] = {(float)(1.0 / 6.0), (float)(1.0 / 6.0), (float)(2.0 / 3.0), 0.0f};
            const float F[4] = {wp.k1[o], wp.k2[o], f, 0.0f};
            float v = y;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float gw = __fmul_rn(b1[q], ik);
              gw = __fadd_rn(gw, __fdiv_rn(__fmul_rn(b2[q], ikk), sq));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b3[q], ik0), rdt));
              gw = __fadd_rn(gw, __fmul_rn(__fmul_rn(b4[q], ikkk), rdt));
              v = __fadd_rn(__fadd_rn(v, __fmul_rn(__fmul_rn(al[q], F[q]), h)), __fmul_rn(sde.gval[stp][q], gw));
            }
            yprev[o] = y; wp.y[o] = v; wp.xs[o] = v;
          }
        }
        __syncwarp();
        if (sg == per - 1) {
          while (emit < sde.n_emit && sde.emit_step[emit] == stp) {
            const float w0 = sde.emit_w0[emit], w1 = sde.emit_w1[emit];
            const size_t ooff = (size_t)sde.emit_out[emit] * P.N * Dd;
            for (int o = lane; o < 4 * Dd; o += 32) {
              const long row = base + o / Dd;
              if (row < P.N) P.out[ooff + row * Dd + (o % Dd)] = __fadd_rn(__fmul_rn(w0, yprev[o]), __fmul_rn(w1, wp.y[o]));
            }
            ++emit;
          }
        }
      }
      __syncwarp();
    }
    __syncwarp();
  }
}

}  // namespace fz

// ---- host side ---------------------------------------------------------------------------------------------------------
struct FusedPlan {
  int W, nj, dt;   // dt = compile-time D (2) or 0 for the wide-D path
  bool stash_smem;
  size_t smem;
};

static bool plan_fused(const idff_weights_t* w, const idff_field_params_t* p, FusedPlan* pl) {
  const NetView& v = w->v;
  if (!(v.w == 128 || v.w == 256) || v.fwt1 == nullptr) return false;
  if (!(p->variant == IDFF_FIELD_MOMENTUM || p->variant == IDFF_FIELD_MD || p->variant == IDFF_FIELD_CFM)) return false;
  if (v.din != p->dim + 1 || v.dout != p->dim) return false;
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  const int dt = p->dim == 2 ? 2 : 0;
  if (!((v.w == 256 && dt == 2) || (v.w == 128 && dt == 0 && nj <= 2))) return false;   // instantiated shapes
  if (p->dim > 64) return false;
  bool stash_smem = true;
  fz::Layout L = fz::make_layout(v.w, nj, p->dim, v.din_pad, v.dout, dt > 0, true);
  if ((size_t)L.total * 4 > 227 * 1024) {
    stash_smem = false;
    L = fz::make_layout(v.w, nj, p->dim, v.din_pad, v.dout, dt > 0, false);
    if ((size_t)L.total * 4 > 227 * 1024) return false;
  }
  pl->W = v.w; pl->nj = nj; pl->dt = dt; pl->stash_smem = stash_smem; pl->smem = (size_t)L.total * 4;
  return true;
}

bool fused_supports(const idff_weights_t* w, const idff_field_params_t* p) {
  FusedPlan pl;
  return plan_fused(w, p, &pl);
}

static int launch_fused(const idff_weights_t* w, const idff_field_params_t* p, fz::FParams& P, const Rk4Table* rk,
                        const SdeTable* sde, const Stage* one, void* stream) {
  FusedPlan pl;
  if (!plan_fused(w, p, &pl)) {
    set_error("fused kernel: unsupported shape/variant (width %d, dim %d)", w->v.w, p->dim);
    return IDFF_EUNSUPPORTED;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, w->device);
  const long tiles = (P.N + fz::TILE - 1) / fz::TILE;
  const int grid = (int)(tiles < sms ? tiles : sms);
  P.stash_g = nullptr;
  if (!pl.stash_smem) {
    idff_weights_t* wm = const_cast<idff_weights_t*>(w);
    const size_t need = (size_t)sms * pl.nj * (pl.W / 32) * 256 * sizeof(float4);
    if (wm->stash_bytes < need) {
      if (wm->d_stash) cudaFree(wm->d_stash);
      wm->d_stash = nullptr;
      wm->stash_bytes = 0;
      IDFF_CUDA(cudaMalloc(&wm->d_stash, need));
      wm->stash_bytes = need;
    }
    P.stash_g = reinterpret_cast<float4*>(wm->d_stash);
  }
  static thread_local Rk4Table rk0;
  static thread_local SdeTable sde0;
  static thread_local Stage one0;
  const Rk4Table& rkr = rk ? *rk : rk0;
  const SdeTable& sder = sde ? *sde : sde0;
  const Stage& oner = one ? *one : one0;
  auto go = [&](auto kern) -> int {
    IDFF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    kern<<<grid, fz::NTHREADS, pl.smem, (cudaStream_t)stream>>>(P, rkr, sder, oner);
    IDFF_CUDA(cudaGetLastError());
    count_launch();
    return IDFF_OK;
  };
  switch (pl.W * 100 + pl.nj * 10 + pl.dt) {
    case 25612: return go(fz::k_fused<256, 1, 2>);
    case 25622: return go(fz::k_fused<256, 2, 2>);
    case 25632: return go(fz::k_fused<256, 3, 2>);
    case 12810: return go(fz::k_fused<128, 1, 0>);
    case 12820: return go(fz::k_fused<128, 2, 0>);
    default:
      set_error("fused kernel: no instance for width %d jets %d dim-mode %d", pl.W, pl.nj, pl.dt);
      return IDFF_EUNSUPPORTED;
  }
}

static void fill_params(const idff_weights_t* w, const idff_field_params_t* p, int mode, int64_t N, fz::FParams* P) {
  memset(P, 0, sizeof(*P));
  int nj, rev;
  jets_needed(*p, &nj, &rev);
  P->net = w->v;
  P->variant = p->variant; P->D = p->dim; P->tidx = p->t_idx; P->reverse = rev; P->mode = mode;
  P->N = (long)N;
}

int fused_field_eval(const idff_weights_t* w, const idff_field_params_t* p, float t, const float* d_x,
                     const float* d_x0, float* d_out, int64_t N, void* stream) {
  fz::FParams P;
  fill_params(w, p, 0, N, &P);
  P.x_in = d_x; P.x0_cfm = d_x0; P.out = d_out;
  Stage st;
  make_stage(*p, t, &st);
  return launch_fused(w, p, P, nullptr, nullptr, &st, stream);
}

int fused_sample_rk4(const idff_weights_t* w, const idff_field_params_t* p, const float* d_x0, const Stage* stages,
                     const float* dts, int n_iv, float* d_out_final, float* d_out_traj, int64_t N, void* stream) {
  static thread_local Rk4Table tab;
  const float* src = d_x0;
  for (int done = 0; done < n_iv; done += kMaxIntervals) {
    const int cnt = n_iv - done < kMaxIntervals ? n_iv - done : kMaxIntervals;
    tab.n_iv = cnt;
    memcpy(tab.dt, dts + done, sizeof(float) * cnt);
    memcpy(tab.st, stages + 4 * (size_t)done, sizeof(Stage) * 4 * cnt);
    fz::FParams P;
    fill_params(w, p, 1, N, &P);
    P.x_in = src; P.out = d_out_final;
    P.x0_cfm = done > 0 ? d_x0 : nullptr;   // CFMModel keeps the x captured at t == 0 for the whole solve
    P.traj = d_out_traj ? d_out_traj + (size_t)done * N * p->dim : nullptr;
    int rc = launch_fused(w, p, P, &tab, nullptr, nullptr, stream);
    if (rc) return rc;
    src = d_out_final;
  }
  return IDFF_OK;
}

int fused_sample_sde(const idff_weights_t* w, const idff_field_params_t* p, const SdeTable& tab, const float* d_y0,
                     const float* d_Ik, const float* d_Ik0, float* d_out, int64_t N, void* stream) {
  fz::FParams P;
  fill_params(w, p, 2, N, &P);
  P.x_in = d_y0; P.out = d_out; P.Ik = d_Ik; P.Ik0 = d_Ik0;
  return launch_fused(w, p, P, nullptr, &tab, nullptr, stream);
}

}  // namespace idff
