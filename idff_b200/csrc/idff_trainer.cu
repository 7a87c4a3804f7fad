// The 2-D IDFF training step as ONE persistent cooperative kernel.
//
// Reference step (2D_examples/Autograd_example_order2.py:89-112, identical in Autograd_example_order3.py:90-109):
//   t, xt, ut, eps = FM.sample_location_and_conditional_flow(x0, x1)        conditional_flow_matching.py:153-193
//   vt   = base_model(cat[xt, t])                                           MLP 3-256-256-256-2, models.py:4-21
//   loss = mean(((1/(1e-2+1-t)) * (vt - FM.x1))**2)                         :105
//   loss.backward(); clip_grad_norm_(params, 1); AdamW(lr=2e-4).step()      :109-112
//
// At the reference's batch (256) this is 0.2 GFLOP: launch- and latency-bound on any GPU (60 torch kernels, 290 us even as
// a CUDA graph).  Here `n_steps` consecutive steps run inside one launch; a step is three phases separated by a grid
// barrier (monotonic counter, one red.release per CTA; the launch is cooperative so all CTAs are co-resident):
//   A  rows:    B/4 CTAs own 4 batch rows each and run path sample -> forward -> loss -> backward-data for their rows
//               with no cross-CTA traffic.  The four 256x256 weight matrices of the step (W1^T, W2^T, W2, W1: 1 MB, L2
//               resident) stream through a 4 x 32 KB shared-memory ring filled by a producer warp with TMA bulk copies
//               (cp.async.bulk + mbarrier complete_tx), running ahead across layer boundaries; a consumer thread owns
//               4 units x 4 rows over one eighth of K (128-bit conflict-free weight reads, broadcast activation reads),
//               the 8 K-slices are summed in a fixed order.  Activations and the pre-activation gradients dZ go to global
//               memory (L2-resident, 1.5 MB).  The producer warp also prefetches the next step's input rows.
//   B  weights: 145 jobs, one per CTA: 2 x 64 tiles (32 x 32) of dW1 / dW2 = dZ^T A (K = batch, operands staged in shared
//               memory, 4 x 4 outputs per thread over one eighth of K), 8 + 8 jobs for the thin layers, one for db3 and the
//               loss value.  A gradient never leaves the CTA that owns its parameter; only the CTA's sum of squares goes
//               to global memory.  The owner's P / exp_avg / exp_avg_sq loads are issued before the barrier.
//   C  update:  every CTA adds the partial sums of squares in the same fixed order (-> identical clip coefficient), then
//               each thread applies clip + AdamW (torch's op order) to the parameters it owns; the transposed forward
//               copies of W1 / W2 are refreshed through a shared-memory transpose (coalesced stores).
// Every reduction has a fixed order: results are bit-reproducible run to run.
#include <cmath>

#include "idff_common.cuh"

namespace idff {
namespace trn {

constexpr int W = 256, DIN = 3, DOUT = 2, R = 4, NC = 512, NT = NC + 32, MAXB = IDFF_TRAIN_MAX_BATCH;
constexpr int oW0 = 0, ob0 = oW0 + W * DIN, oW1 = ob0 + W, ob1 = oW1 + W * W, oW2 = ob1 + W, ob2 = oW2 + W * W,
              oW3 = ob2 + W, ob3 = oW3 + DOUT * W, NP = ob3 + DOUT;
constexpr int NPP = (NP + 63) / 64 * 64;
constexpr int JOB_W0 = 128, JOB_W3 = 136, JOB_B3 = 144, N_JOBS = 145;
constexpr int MAX_GRID = 256;
constexpr int NS = 4, CH_ROWS = 32, CH_FLOATS = CH_ROWS * W, CH_BYTES = CH_FLOATS * 4, CH_PER_PASS = W / CH_ROWS,
              CH_PER_STEP = 4 * CH_PER_PASS;
static_assert(CH_PER_STEP % NS == 0, "ring phase bookkeeping");

// clock64 phase stamps of the last step of the last launch, per CTA: diagnostic only (idff_debug_train_stamps)
__device__ long long g_train_stamps[256][8];

// offsets (floats) into the caller-owned state buffer
constexpr size_t sP = 0, sM = sP + NPP, sV = sM + NPP, sWT1 = sV + NPP, sWT2 = sWT1 + W * W, sIN = sWT2 + W * W,
                 sA1 = sIN + (size_t)MAXB * 4, sA2 = sA1 + (size_t)MAXB * W, sA3 = sA2 + (size_t)MAXB * W,
                 sZ1 = sA3 + (size_t)MAXB * W, sZ2 = sZ1 + (size_t)MAXB * W, sZ3 = sZ2 + (size_t)MAXB * W,
                 sDO = sZ3 + (size_t)MAXB * W, sLP = sDO + (size_t)MAXB * DOUT, sNP = sLP + MAX_GRID, sBAR = sNP + MAX_GRID,
                 sTOTAL = sBAR + 64;

// dynamic shared memory (floats).  Phase A: ring | act | part | raw | in4 | dout | lterm.  Phases B / C alias it.
constexpr int mRING = 0, mACT = mRING + NS * CH_FLOATS, mPART = mACT + R * W, mRAW = mPART + 8 * R * W, mIN4 = mRAW + 2 * R * 8,
              mDOUT = mIN4 + R * 4, mLT = mDOUT + R * DOUT, mTOTAL = mLT + 8;
constexpr int mDZS = 0, mAS = mDZS + 256 * 32, mRED = mAS + 256 * 32, mREDB = mRED + 8 * 16 * 64, mPT = mREDB + 16 * 32,
              mBTOTAL = mPT + 32 * 33;
static_assert(mBTOTAL <= mRAW, "phase B scratch must not overlap the input prefetch buffer");
constexpr size_t kSmemBytes = (size_t)mTOTAL * sizeof(float);

struct Args {
  float* S;   // state
  const float *x0, *x1, *t, *eps;
  float *loss_out, *gnorm_out, *grad_dump;
  int B, n_steps;
  long long steps_done;
  float sigma;
  int sigma_mode;
  double lr, beta1, beta2, b1pow, b2pow;            // bias corrections in double like torch's Python scalars; b?pow = beta?^steps_done
  float decay, omb1, omb2, beta2f, eps_adam, clip;  // 1 - lr*wd, 1 - beta1, 1 - beta2 (rounded to fp32 from double)
};

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(ok)
        : "r"(s32(b)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(dst)),
               "l"(src), "r"(bytes), "r"(s32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void csync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }      // the 16 consumer warps
__device__ __forceinline__ void fullsync() { asm volatile("bar.sync 0, 544;" ::: "memory"); }   // + the producer warp

__device__ __forceinline__ float sigma_t_of(float t, float sigma, int mode) {
  if (mode == 0) return sigma;
  return __fmul_rn(sigma, __fsqrt_rn(__fmul_rn(t, __fsub_rn(1.0f, t))));   // conditional_flow_matching.py:231-233
}

// Grid barrier for the consumer warps: one release-add per CTA on a monotonic counter, then poll.  `with_producer`: the
// closing CTA barrier also admits the producer warp (top of the next step).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned& epoch, bool with_producer) {
  csync();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ctr), "r"(1u) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < epoch);
  }
  if (with_producer) fullsync(); else csync();
}

__device__ __forceinline__ float selu_d1(float code) { return code < 0.0f ? kSeluLambda : code; }

// One 256 x 256 layer pass over the ring: thread (ks, u4) accumulates units 4 u4 .. 4 u4 + 3 of the CTA's 4 rows over the
// k rows {32 c + 4 ks + i} of every chunk c, then the 8 K-slices meet in `part`.
__device__ __forceinline__ void gemm_ring(const float* ring, const float* act, float* part, uint64_t* full, uint64_t* empty,
                                          uint32_t& cc, int ks, int u4, int lane) {
  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r) { acc[r][0] = 0.f; acc[r][1] = 0.f; acc[r][2] = 0.f; acc[r][3] = 0.f; }
#pragma unroll 1
  for (int c = 0; c < CH_PER_PASS; ++c) {
    const uint32_t stage = cc % NS;
    mbar_wait(&full[stage], (cc / NS) & 1);
    const float* wb = ring + stage * CH_FLOATS + (4 * ks) * W + 4 * u4;
    float4 av[R];
#pragma unroll
    for (int r = 0; r < R; ++r) av[r] = *reinterpret_cast<const float4*>(act + r * W + c * CH_ROWS + 4 * ks);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(wb + i * W);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float x = i == 0 ? av[r].x : i == 1 ? av[r].y : i == 2 ? av[r].z : av[r].w;
        acc[r][0] = fmaf(x, w.x, acc[r][0]);
        acc[r][1] = fmaf(x, w.y, acc[r][1]);
        acc[r][2] = fmaf(x, w.z, acc[r][2]);
        acc[r][3] = fmaf(x, w.w, acc[r][3]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[stage]);
    ++cc;
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
    *reinterpret_cast<float4*>(part + (ks * R + r) * W + 4 * u4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
}

__device__ __forceinline__ float part_sum(const float* part, int r, int j) {
  float z = part[r * W + j];
#pragma unroll
  for (int k = 1; k < 8; ++k) z += part[(k * R + r) * W + j];
  return z;
}

// raw input rows of one step -> shared memory: {x0[0], x0[1], x1[0], x1[1], eps[0], eps[1], t, 0}
__device__ __forceinline__ void load_raw(const Args& a, int s, int row_in_cta, int cta, float* raw) {
  const size_t row = (size_t)s * a.B + cta * R + row_in_cta;
  const float2 x0 = *reinterpret_cast<const float2*>(a.x0 + row * 2), x1 = *reinterpret_cast<const float2*>(a.x1 + row * 2);
  const float2 e = a.eps ? *reinterpret_cast<const float2*>(a.eps + row * 2) : make_float2(0.f, 0.f);
  float* o = raw + ((s & 1) * R + row_in_cta) * 8;
  *reinterpret_cast<float4*>(o) = make_float4(x0.x, x0.y, x1.x, x1.y);
  *reinterpret_cast<float4*>(o + 4) = make_float4(e.x, e.y, a.t[row], 0.f);
}

// ---- phase A (consumer warps): path sample, forward, loss, backward-data for rows [4 cta, 4 cta + 4) ----------------------
__device__ void phase_rows(const Args& a, int s, float* smem, int cta, uint64_t* full, uint64_t* empty, uint32_t& cc) {
  const float* ring = smem + mRING;
  float* act = smem + mACT;       // [R][W]
  float* part = smem + mPART;     // [8][R][W]
  const float* raw = smem + mRAW + (s & 1) * R * 8;
  float* in4 = smem + mIN4;       // [R][4]   xt0, xt1, t, 1
  float* dout = smem + mDOUT;     // [R][2]
  float* lterm = smem + mLT;      // [R*DOUT]
  const int tid = threadIdx.x, h = tid >> 8, j = tid & 255, lane = tid & 31, warp = tid >> 5;
  const int ks = tid >> 6, u4 = tid & 63;
  const int row0 = cta * R;
  const float* P = a.S + sP;

  // this thread's share of the thin layers and biases: issued now, needed later
  const float w00 = __ldcg(P + oW0 + j * 3), w01 = __ldcg(P + oW0 + j * 3 + 1), w02 = __ldcg(P + oW0 + j * 3 + 2);
  const float bias0 = __ldcg(P + ob0 + j), bias1 = __ldcg(P + ob1 + j), bias2 = __ldcg(P + ob2 + j);
  const float w30 = __ldcg(P + oW3 + j), w31 = __ldcg(P + oW3 + W + j);

  if (tid < R) {   // K2 inlined: xt = t*x1 + (1-t)*x0 + sigma_t*eps   (conditional_flow_matching.py:98-123, no contraction)
    const float4 q0 = *reinterpret_cast<const float4*>(raw + tid * 8), q1 = *reinterpret_cast<const float4*>(raw + tid * 8 + 4);
    const float t = q1.z;
    const float st = sigma_t_of(t, a.sigma, a.sigma_mode);
    const float omt = __fsub_rn(1.0f, t);
    const float xt0 = __fadd_rn(__fadd_rn(__fmul_rn(t, q0.z), __fmul_rn(omt, q0.x)), __fmul_rn(st, q1.x));
    const float xt1 = __fadd_rn(__fadd_rn(__fmul_rn(t, q0.w), __fmul_rn(omt, q0.y)), __fmul_rn(st, q1.y));
    const float4 v = make_float4(xt0, xt1, t, 1.0f);
    *reinterpret_cast<float4*>(in4 + tid * 4) = v;
    *reinterpret_cast<float4*>(a.S + sIN + (size_t)(row0 + tid) * 4) = v;
  }
  csync();

  float code1[2], code2[2], code3[2];
#pragma unroll
  for (int q = 0; q < 2; ++q) {   // layer 1 (K = 3)
    const int r = 2 * h + q;
    const float z = fmaf(in4[r * 4 + 2], w02, fmaf(in4[r * 4 + 1], w01, fmaf(in4[r * 4], w00, bias0)));
    const float av = selu_fwd(z, code1[q]);
    act[r * W + j] = av;
    a.S[sA1 + (size_t)(row0 + r) * W + j] = av;
  }
  csync();

#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // layers 2, 3: ring passes 0 (W1^T), 1 (W2^T)
    gemm_ring(ring, act, part, full, empty, cc, ks, u4, lane);
    csync();
    const float b = l ? bias2 : bias1;
    float* Aout = a.S + (l ? sA3 : sA2);
    float avs[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      const float z = part_sum(part, r, j) + b;
      float c;
      avs[q] = selu_fwd(z, c);
      if (l) code3[q] = c; else code2[q] = c;
      Aout[(size_t)(row0 + r) * W + j] = avs[q];
    }
    act[(2 * h) * W + j] = avs[0];       // every gemm_ring read of act happened before the csync above
    act[(2 * h + 1) * W + j] = avs[1];
    csync();
  }

  // output layer + loss (Autograd_example_order2.py:105) + d loss / d vt
  if (warp < R * DOUT) {
    const int r = warp >> 1, o = warp & 1;
    float sum = 0.0f;
#pragma unroll
    for (int i = 0; i < W / 32; ++i) sum = fmaf(act[r * W + lane + 32 * i], __ldcg(P + oW3 + o * W + lane + 32 * i), sum);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (lane == 0) {
      const float vt = sum + __ldcg(P + ob3 + o);
      const float t = in4[r * 4 + 2];
      const float x1v = raw[r * 8 + 2 + o];
      const float wgt = __fdiv_rn(1.0f, __fsub_rn(1.01f, t));
      const float rr = __fmul_rn(wgt, __fsub_rn(vt, x1v));
      lterm[warp] = __fmul_rn(rr, rr);
      const float gk = 1.0f / (float)(a.B * DOUT);
      const float d = 2.0f * rr * wgt * gk;
      dout[r * DOUT + o] = d;
      a.S[sDO + (size_t)(row0 + r) * DOUT + o] = d;
    }
  }
  csync();
  if (tid == 0) {
    float ls = 0.0f;
#pragma unroll
    for (int i = 0; i < R * DOUT; ++i) ls += lterm[i];
    a.S[sLP + cta] = ls;
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {   // dZ3 = (dOut W3) * selu'(z3)
    const int r = 2 * h + q;
    const float dz = fmaf(dout[r * 2 + 1], w31, dout[r * 2] * w30) * selu_d1(code3[q]);
    act[r * W + j] = dz;
    a.S[sZ3 + (size_t)(row0 + r) * W + j] = dz;
  }
  csync();
#pragma unroll 1
  for (int l = 0; l < 2; ++l) {   // dZ2 = (dZ3 W2) * selu'(z2), dZ1 = (dZ2 W1) * selu'(z1): ring passes 2 (W2), 3 (W1)
    gemm_ring(ring, act, part, full, empty, cc, ks, u4, lane);
    csync();
    float* Zout = a.S + (l ? sZ1 : sZ2);
    float dzs[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int r = 2 * h + q;
      dzs[q] = part_sum(part, r, j) * selu_d1(l ? code1[q] : code2[q]);
      Zout[(size_t)(row0 + r) * W + j] = dzs[q];
    }
    act[(2 * h) * W + j] = dzs[0];
    act[(2 * h + 1) * W + j] = dzs[1];
    csync();
  }
}

// One thread's share of the parameters: up to 4 gradients and their flat indices (-1: none).
struct Owned {
  float g[4];
  int idx[4];
  int wt_off;   // tile jobs: state offset of WT[kb*32][jb*32] (the 32 x 32 transposed block of this CTA), else -1
};

// ---- phase B ---------------------------------------------------------------------------------------------------------------
// tile (jb, kb) of dW = dZ^T A: thread (ksl, jj, kk) accumulates the 4 x 4 block (4 jj.., 4 kk..) over rows r = ksl (mod 8);
// the 8 partial blocks meet in shared memory; thread (j = tid / 16, kp = tid % 16) ends up owning (j, 2 kp), (j, 2 kp + 1).
__device__ void job_tile(const Args& a, int job, float* smem, Owned& own) {
  float* dZs = smem + mDZS;     // [256][32]
  float* As = smem + mAS;       // [256][32]
  float* red = smem + mRED;     // [8][16][64]
  float* redb = smem + mREDB;   // [16][32]
  const int tid = threadIdx.x;
  const int m = job >> 6, tile = job & 63, jb = tile >> 3, kb = tile & 7;
  const float* dZ = a.S + (m ? sZ3 : sZ2);
  const float* Ain = a.S + (m ? sA2 : sA1);
  const int ksl = tid >> 6, t64 = tid & 63, jj = t64 >> 3, kk = t64 & 7;
  float acc[4][4], gbs = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
  for (int r0 = 0; r0 < a.B; r0 += 256) {
    const int rows = min(256, a.B - r0);
    {   // stage the two 256 x 32 operand panels: all 8 128-bit loads of a thread in flight at once
      float4 vz[4], va[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + u * NC, r = i >> 3, c4 = i & 7;
        if (r < rows) {
          vz[u] = __ldcg(reinterpret_cast<const float4*>(dZ + (size_t)(r0 + r) * W + jb * 32) + c4);
          va[u] = __ldcg(reinterpret_cast<const float4*>(Ain + (size_t)(r0 + r) * W + kb * 32) + c4);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = tid + u * NC, r = i >> 3, c4 = i & 7;
        if (r < rows) {
          *reinterpret_cast<float4*>(dZs + r * 32 + c4 * 4) = vz[u];
          *reinterpret_cast<float4*>(As + r * 32 + c4 * 4) = va[u];
        }
      }
    }
    csync();
#pragma unroll 4
    for (int r = ksl; r < rows; r += 8) {
      const float4 dz = *reinterpret_cast<const float4*>(dZs + r * 32 + 4 * jj);
      const float4 av = *reinterpret_cast<const float4*>(As + r * 32 + 4 * kk);
      const float d[4] = {dz.x, dz.y, dz.z, dz.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(d[i], av.x, acc[i][0]);
        acc[i][1] = fmaf(d[i], av.y, acc[i][1]);
        acc[i][2] = fmaf(d[i], av.z, acc[i][2]);
        acc[i][3] = fmaf(d[i], av.w, acc[i][3]);
      }
    }
    if (kb == 0) {   // column sums of dZ: the bias gradient (tiles of the first k-block only); thread = (16 row groups, column)
      const int c = tid & 31, rg = tid >> 5;
      float sum = 0.f;
      for (int r = rg; r < rows; r += 16) sum += dZs[r * 32 + c];
      gbs += sum;
    }
    csync();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) red[(ksl * 16 + i * 4 + c) * 64 + t64] = acc[i][c];
  if (kb == 0) redb[tid] = gbs;   // [16 row groups][32 columns]
  csync();
  {
    const int j = tid >> 4, kp = tid & 15;
    const int src = (j >> 2) * 8 + (kp >> 1), e = (j & 3) * 4 + (kp & 1) * 2;
    float g0 = 0.f, g1 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      g0 += red[(k * 16 + e) * 64 + src];
      g1 += red[(k * 16 + e + 1) * 64 + src];
    }
    const int oW = m ? oW2 : oW1;
    own.g[0] = g0; own.g[1] = g1;
    own.idx[0] = oW + (jb * 32 + j) * W + kb * 32 + 2 * kp;
    own.idx[1] = own.idx[0] + 1;
    own.wt_off = (int)(m ? sWT2 : sWT1) + (kb * 32) * W + jb * 32;
    if (kb == 0 && kp == 0) {
      float b = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) b += redb[k * 32 + j];
      own.g[2] = b;
      own.idx[2] = (m ? ob2 : ob1) + jb * 32 + j;
    }
  }
}

// thin layers: G[j][c] = sum_r X[r][j] * Y[r][c]; X = dZ1, Y = IN (c < 4, IN[r][3] = 1 -> db0) or X = A3, Y = dOut (c < 2)
template <int C>
__device__ void job_thin(const Args& a, const float* X, const float* Y, int q, float* smem, float (&out)[4]) {
  const int tid = threadIdx.x, jl = tid & 31, rg = tid >> 5, j = 32 * q + jl;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
  for (int r = rg; r < a.B; r += 16) {
    const float x = __ldcg(X + (size_t)r * W + j);
    if (C == 4) {
      const float4 y = __ldcg(reinterpret_cast<const float4*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]); acc[2] = fmaf(x, y.z, acc[2]); acc[3] = fmaf(x, y.w, acc[3]);
    } else {
      const float2 y = __ldcg(reinterpret_cast<const float2*>(Y) + r);
      acc[0] = fmaf(x, y.x, acc[0]); acc[1] = fmaf(x, y.y, acc[1]);
    }
  }
  *reinterpret_cast<float4*>(smem + (rg * 32 + jl) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  csync();
  if (tid < 32) {
#pragma unroll
    for (int c = 0; c < 4; ++c) out[c] = 0.0f;
#pragma unroll
    for (int g = 0; g < 16; ++g) {
      const float4 v = *reinterpret_cast<const float4*>(smem + (g * 32 + tid) * 4);
      out[0] += v.x; out[1] += v.y; out[2] += v.z; out[3] += v.w;
    }
  }
}

__global__ void __launch_bounds__(NT, 1) k_train_flow(const Args a) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t full[NS], empty[NS];
  __shared__ double sred[NC / 32];
  __shared__ float s_coef, s_step_size, s_bc2_sqrt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cta = blockIdx.x;
  const int nA = a.B / R;
  const bool row_cta = cta < nA;
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NC / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (row_cta && tid < R) load_raw(a, 0, tid, cta, smem + mRAW);
  __syncthreads();
  uint32_t cc = 0;   // ring chunks issued (producer) / consumed (consumers): the same sequence on both sides

  if (warp == NC / 32) {
    // ---- producer warp: streams the step's four weight matrices through the ring, prefetches the next step's inputs ----
    for (int s = 0; s < a.n_steps; ++s) {
      if (s > 0) fullsync();   // weights of step s are final (the consumers passed the closing grid barrier of step s - 1)
      if (!row_cta) continue;
      if (lane == 0) {
        fence_async_global();
        const float* src[4] = {a.S + sWT1, a.S + sWT2, a.S + sP + oW2, a.S + sP + oW1};
#pragma unroll 1
        for (int c = 0; c < CH_PER_STEP; ++c, ++cc) {
          const uint32_t stage = cc % NS;
          if (cc >= NS) mbar_wait(&empty[stage], ((cc / NS) - 1) & 1);
          mbar_arrive_expect_tx(&full[stage], CH_BYTES);
          tma_load_1d(smem + mRING + stage * CH_FLOATS, src[c / CH_PER_PASS] + (size_t)(c % CH_PER_PASS) * CH_FLOATS, CH_BYTES,
                      &full[stage]);
        }
      }
      if (s + 1 < a.n_steps && lane < R) load_raw(a, s + 1, lane, cta, smem + mRAW);
    }
    return;
  }

  // ---- consumer warps ------------------------------------------------------------------------------------------------------
  unsigned epoch = 0;
  unsigned* bar = reinterpret_cast<unsigned*>(a.S + sBAR);
  float* P = a.S + sP;
  float* M = a.S + sM;
  float* V = a.S + sV;
  double b1pow = a.b1pow, b2pow = a.b2pow;

  for (int s = 0; s < a.n_steps; ++s) {
    const bool stamp = tid == 0 && s == a.n_steps - 1;
    if (stamp) g_train_stamps[cta][0] = clock64();
    if (row_cta) phase_rows(a, s, smem, cta, full, empty, cc);
    if (stamp) g_train_stamps[cta][1] = clock64();
    grid_barrier(bar, epoch, false);
    if (stamp) g_train_stamps[cta][2] = clock64();

    Owned own;
#pragma unroll
    for (int i = 0; i < 4; ++i) { own.g[i] = 0.0f; own.idx[i] = -1; }
    own.wt_off = -1;
    if (cta < JOB_W0) {
      job_tile(a, cta, smem, own);
    } else if (cta < JOB_W3) {   // dW0, db0 for units [32 q, 32 q + 32)
      float o[4];
      const int q = cta - JOB_W0;
      job_thin<4>(a, a.S + sZ1, a.S + sIN, q, smem, o);
      if (tid < 32) {
        const int j = 32 * q + tid;
#pragma unroll
        for (int c = 0; c < 3; ++c) { own.g[c] = o[c]; own.idx[c] = oW0 + j * 3 + c; }
        own.g[3] = o[3]; own.idx[3] = ob0 + j;
      }
    } else if (cta < JOB_B3) {   // dW3 for units [32 q, 32 q + 32)
      float o[4];
      const int q = cta - JOB_W3;
      job_thin<2>(a, a.S + sA3, a.S + sDO, q, smem, o);
      if (tid < 32) {
        const int j = 32 * q + tid;
        own.g[0] = o[0]; own.idx[0] = oW3 + j;
        own.g[1] = o[1]; own.idx[1] = oW3 + W + j;
      }
    } else if (cta == JOB_B3) {   // db3 and the loss value
      if (warp < 2) {
        float sum = 0.0f;
        for (int r = lane; r < a.B; r += 32) sum += __ldcg(a.S + sDO + (size_t)r * DOUT + warp);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0) { own.g[0] = sum; own.idx[0] = ob3 + warp; }
      } else if (warp == 2) {
        double sum = 0.0;
        for (int c = lane; c < nA; c += 32) sum += (double)__ldcg(a.S + sLP + c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        if (lane == 0 && a.loss_out) a.loss_out[s] = (float)(sum / (double)(a.B * DOUT));
      }
    }
    // the owner's parameter and moment loads do not depend on the clip coefficient: in flight across the barrier
    float pv[4], mv[4], vv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int ix = own.idx[i];
      pv[i] = mv[i] = vv[i] = 0.0f;
      if (ix >= 0) { pv[i] = __ldcg(P + ix); mv[i] = __ldcg(M + ix); vv[i] = __ldcg(V + ix); }
    }
    {   // this CTA's sum of squared gradients
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < 4; ++i) ss += (double)own.g[i] * (double)own.g[i];
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
      if (lane == 0) sred[warp] = ss;
      csync();
      if (tid == 0) {
        double tot = 0.0;
        for (int i = 0; i < NC / 32; ++i) tot += sred[i];
        a.S[sNP + cta] = (float)tot;
      }
    }
    if (stamp) g_train_stamps[cta][3] = clock64();
    grid_barrier(bar, epoch, false);
    if (stamp) g_train_stamps[cta][4] = clock64();

    b1pow *= a.beta1;
    b2pow *= a.beta2;
    if (warp == 0) {   // clip_grad_norm_(params, clip): coef = min(1, clip / (norm + 1e-6))
      double tot = 0.0;
#pragma unroll
      for (int i = 0; i < MAX_GRID / 32; ++i) {
        const int c = lane + 32 * i;
        tot += c < (int)gridDim.x ? (double)__ldcg(a.S + sNP + c) : 0.0;
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
      if (lane == 0) {
        const float norm = (float)sqrt(tot);
        s_coef = a.clip > 0.0f ? fminf(1.0f, a.clip / (norm + 1e-6f)) : 1.0f;
        if (cta == 0 && a.gnorm_out) a.gnorm_out[s] = norm;
        s_step_size = (float)(a.lr / (1.0 - b1pow));
        s_bc2_sqrt = (float)sqrt(1.0 - b2pow);
      }
    }
    csync();
    {   // AdamW, torch/optim/adamw.py op order
      const float coef = s_coef, step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
      float pn[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        pn[i] = 0.0f;
        const int ix = own.idx[i];
        if (ix < 0) continue;
        if (a.grad_dump && s == a.n_steps - 1) a.grad_dump[ix] = own.g[i];
        const float g = own.g[i] * coef;
        const float p = pv[i] * a.decay;
        const float mm = fmaf(g - mv[i], a.omb1, mv[i]);
        const float v2 = fmaf(a.omb2 * g, g, vv[i] * a.beta2f);
        const float denom = sqrtf(v2) / bc2_sqrt + a.eps_adam;
        pn[i] = p - step_size * (mm / denom);
        P[ix] = pn[i]; M[ix] = mm; V[ix] = v2;
      }
      if (own.wt_off >= 0) {   // W^T copy of the tile through shared memory: rows of 32 consecutive j per k
        float* pt = smem + mPT;
        const int j = tid >> 4, kp = tid & 15;
        pt[(2 * kp) * 33 + j] = pn[0];
        pt[(2 * kp + 1) * 33 + j] = pn[1];
        csync();
        const int k = tid >> 4, j2 = (tid & 15) * 2;
        *reinterpret_cast<float2*>(a.S + own.wt_off + k * W + j2) = make_float2(pt[k * 33 + j2], pt[k * 33 + j2 + 1]);
      }
      fence_async_global();   // the next step's TMA reads (async proxy) of these generic-proxy writes
    }
    if (stamp) g_train_stamps[cta][5] = clock64();
    grid_barrier(bar, epoch, s + 1 < a.n_steps);
    if (stamp) g_train_stamps[cta][6] = clock64();
  }
}

__global__ void k_train_init(float* S) {   // WT[k][j] = W[j][k] for layers 2 and 3
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < W * W) {
    const int k = i / W, j = i % W;
    S[sWT1 + i] = S[sP + oW1 + j * W + k];
    S[sWT2 + i] = S[sP + oW2 + j * W + k];
  }
}

}  // namespace trn
}  // namespace idff

using namespace idff;
using namespace idff::trn;

extern "C" int64_t idff_train_flow_state_floats(void) { return (int64_t)sTOTAL; }

extern "C" int idff_train_flow_offsets(int64_t* h_out12) {
  IDFF_CHECK_ARG(h_out12, "idff_train_flow_offsets: null pointer");
  const int64_t v[12] = {oW0, ob0, oW1, ob1, oW2, ob2, oW3, ob3, NP, (int64_t)sP, (int64_t)sM, (int64_t)sV};
  for (int i = 0; i < 12; ++i) h_out12[i] = v[i];
  return IDFF_OK;
}

extern "C" int idff_debug_train_stamps(int64_t* h_out, int32_t n_ctas) {
  IDFF_CHECK_ARG(h_out && n_ctas >= 1 && n_ctas <= 256, "idff_debug_train_stamps: bad argument");
  IDFF_CUDA(cudaMemcpyFromSymbol(h_out, g_train_stamps, sizeof(long long) * 8 * n_ctas));
  return IDFF_OK;
}

extern "C" int idff_train_flow_init(float* d_state, void* stream) {
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_flow_init: state must be a 16-byte aligned device pointer");
  k_train_init<<<(W * W + 255) / 256, 256, 0, (cudaStream_t)stream>>>(d_state);
  IDFF_CUDA(cudaGetLastError());
  count_launch();
  return IDFF_OK;
}

extern "C" int idff_train_flow_steps(float* d_state, const float* d_x0, const float* d_x1, const float* d_t,
                                     const float* d_eps, int32_t n_steps, int32_t B, float sigma, int32_t sigma_mode,
                                     const idff_adamw_t* opt, int64_t steps_done, float* d_loss, float* d_gnorm,
                                     float* d_grad_dump, void* stream) {
  IDFF_CHECK_ARG(d_state && (reinterpret_cast<uintptr_t>(d_state) & 15) == 0, "idff_train_flow_steps: state must be a 16-byte aligned device pointer");
  IDFF_CHECK_ARG(d_x0 && d_x1 && d_t && opt, "idff_train_flow_steps: null pointer");
  IDFF_CHECK_ARG(((reinterpret_cast<uintptr_t>(d_x0) | reinterpret_cast<uintptr_t>(d_x1) | reinterpret_cast<uintptr_t>(d_eps)) & 7) == 0,
                 "idff_train_flow_steps: x0, x1, eps must be 8-byte aligned");
  IDFF_CHECK_ARG(n_steps >= 0 && n_steps <= (1 << 20), "idff_train_flow_steps: n_steps %d outside [0, 2^20]", n_steps);
  IDFF_CHECK_ARG(B >= R && B <= MAXB && B % R == 0, "idff_train_flow_steps: batch %d must be a multiple of %d in [%d, %d]", B, R, R, MAXB);
  IDFF_CHECK_ARG(sigma_mode == 0 || sigma_mode == 1, "idff_train_flow_steps: sigma_mode %d", sigma_mode);
  IDFF_CHECK_ARG(steps_done >= 0, "idff_train_flow_steps: steps_done < 0");
  if (n_steps == 0) return IDFF_OK;
  int dev = 0, sms = 0, coop = 0;
  IDFF_CUDA(cudaGetDevice(&dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  IDFF_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  if (!coop || sms < N_JOBS) {
    set_error("idff_train_flow_steps: needs cooperative launch and >= %d SMs (device has %d)", N_JOBS, sms);
    return IDFF_EUNSUPPORTED;
  }
  const int grid = sms < MAX_GRID ? sms : MAX_GRID;
  static bool attr_set = false;
  if (!attr_set) {
    IDFF_CUDA(cudaFuncSetAttribute(k_train_flow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    attr_set = true;
  }
  Args a;
  a.S = d_state;
  a.x0 = d_x0; a.x1 = d_x1; a.t = d_t; a.eps = d_eps;
  a.loss_out = d_loss; a.gnorm_out = d_gnorm; a.grad_dump = d_grad_dump;
  a.B = B; a.n_steps = n_steps; a.steps_done = steps_done;
  a.sigma = sigma; a.sigma_mode = sigma_mode;
  a.lr = opt->lr; a.beta1 = opt->beta1; a.beta2 = opt->beta2;
  a.decay = (float)(1.0 - opt->lr * opt->weight_decay);
  a.omb1 = (float)(1.0 - opt->beta1); a.omb2 = (float)(1.0 - opt->beta2);
  a.beta2f = (float)opt->beta2; a.eps_adam = (float)opt->eps; a.clip = (float)opt->max_grad_norm;
  a.b1pow = std::pow(opt->beta1, (double)steps_done); a.b2pow = std::pow(opt->beta2, (double)steps_done);
  cudaStream_t st = (cudaStream_t)stream;
  IDFF_CUDA(cudaMemsetAsync(d_state + sBAR, 0, 64 * sizeof(float), st));
  void* kargs[] = {&a};
  IDFF_CUDA(cudaLaunchCooperativeKernel((const void*)k_train_flow, dim3(grid), dim3(NT), kargs, kSmemBytes, st));
  count_launch();
  return IDFF_OK;
}
