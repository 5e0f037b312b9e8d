// K4: utility-weighted column reductions over the population (one fused pass, two accumulators per column).
//   S1_j = sum_r a_r eps_rj          S2_j = sum_r b_r (eps_rj^2 * c1_j - c0_j)         eps = X - mu
// with (c1, c0) = (1/sigma, sigma) for the PGPE forms, (1/sigma^2, 1) for SNES, (1, 0) for raw moments.
// HBM-bound: each CTA owns a column tile (128-bit loads, 4 rows in flight per thread) and a contiguous chunk of
// rows; partial sums go to a [chunk][2][D] workspace and a second tiny kernel adds the chunks in a fixed order
// (deterministic, no atomics).  In the symmetric form only the even ("+") rows are read.
#include <cstdlib>

#include "evok_common.cuh"

#ifndef EVOK_GRAD_TMA_DEFAULT
#define EVOK_GRAD_TMA_DEFAULT 1
#endif

namespace evok {

#ifndef EVOK_GRAD_UNROLL
#define EVOK_GRAD_UNROLL 4
#endif
#ifndef EVOK_GRAD_MINB
#define EVOK_GRAD_MINB 4
#endif
#ifndef EVOK_GRAD_CTAS_PER_SM
#define EVOK_GRAD_CTAS_PER_SM 4
#endif
constexpr int kGradThreads = 256;
constexpr int kGradUnroll = EVOK_GRAD_UNROLL;
constexpr int kMaxResidentCtas = 148 * 8;

template <int VEC>
struct VecF;
template <>
struct VecF<4> {
  float v[4];
};
template <>
struct VecF<1> {
  float v[1];
};

template <int VEC>
__device__ __forceinline__ VecF<VEC> load_row(const float* p) {
  VecF<VEC> r;
  if (VEC == 4) {
    const float4 t = ld_stream4(p);
    r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w;
  } else {
    r.v[0] = ld_stream1(p);
  }
  return r;
}

// Batched searches (functional API): blockIdx.z = batch item; element strides between the items' operands (0 = shared)
struct GradItems {
  int64_t x, w, mu, sigma, partial;
};

// SYM: unit r = direction (rows 2r, 2r+1), else unit r = row r.  REGEN: eps = sigma * z regenerated from Philox.
template <int VEC, int TX, bool SYM, bool REGEN>
__global__ void __launch_bounds__(kGradThreads, EVOK_GRAD_MINB)
    grad_partial_kernel(int form, const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, const float* __restrict__ mu,
                        const float* __restrict__ sigma, int64_t n_units, int64_t D, int64_t units_per_chunk, uint64_t unit0,
                        const __grid_constant__ PhiloxKey key, const uint32_t* __restrict__ stream_off, float* __restrict__ partial,
                        const __grid_constant__ GradItems items) {
  constexpr int TY = kGradThreads / TX;
  if (gridDim.z > 1) {
    const int64_t item = blockIdx.z;
    X += item * items.x;
    w += item * items.w;
    mu += item * items.mu;
    sigma += item * items.sigma;
    partial += item * items.partial;
  }
  const uint32_t sw = key.stream_lo + ((REGEN && stream_off) ? __ldg(stream_off) : 0u);
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int64_t col = ((int64_t)blockIdx.x * TX + tx) * VEC;
  const bool active = col < D;

  float m[VEC], c1[VEC], c0[VEC], sg[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) {
    const bool ok = active && (col + c < D);
    const float s = ok ? __ldg(sigma + col + c) : 1.0f;
    m[c] = ok ? __ldg(mu + col + c) : 0.0f;
    sg[c] = s;
    if (form == EVOK_GRAD_EXP) {
      c1[c] = __fdiv_rn(1.0f, s * s);
      c0[c] = 1.0f;
    } else if (form == EVOK_GRAD_MOMENTS) {
      c1[c] = 1.0f;
      c0[c] = 0.0f;
    } else {
      c1[c] = __fdiv_rn(1.0f, s);
      c0[c] = s;
    }
  }

  float s1[VEC], s2[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) s1[c] = s2[c] = 0.0f;

  const int64_t r_begin = (int64_t)blockIdx.y * units_per_chunk;
  const int64_t r_end = min(n_units, r_begin + units_per_chunk);

  for (int64_t r0 = r_begin + ty; r0 < r_end; r0 += (int64_t)TY * kGradUnroll) {
    float a[kGradUnroll], b[kGradUnroll];
    bool need[kGradUnroll];
    VecF<VEC> x[kGradUnroll];
#pragma unroll
    for (int u = 0; u < kGradUnroll; ++u) {
      const int64_t r = r0 + (int64_t)u * TY;
      a[u] = b[u] = 0.0f;
      if (r < r_end) {
        if (SYM) {
          const float wp = __ldg(w + 2 * r), wm = __ldg(w + 2 * r + 1);
          a[u] = 0.5f * (wp - wm);
          b[u] = 0.5f * (wp + wm);
        } else {
          a[u] = b[u] = __ldg(w + r);
        }
      }
      need[u] = active && (a[u] != 0.0f || b[u] != 0.0f);
    }
#pragma unroll
    for (int u = 0; u < kGradUnroll; ++u) {
      const int64_t r = r0 + (int64_t)u * TY;
      if (need[u]) {
        if (REGEN) {
          if (VEC == 4) {
            normals4(key, sw, unit0 + (uint64_t)r, (uint32_t)(col >> 2), x[u].v);
          } else {
            float z[4];
            normals4(key, sw, unit0 + (uint64_t)r, (uint32_t)(col >> 2), z);
            x[u].v[0] = z[col & 3];
          }
        } else {
          x[u] = load_row<VEC>(X + (SYM ? 2 * r : r) * ldx + col);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kGradUnroll; ++u) {
      if (need[u]) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) {
          const float e = REGEN ? sg[c] * x[u].v[c] : x[u].v[c] - m[c];
          s1[c] = fmaf(a[u], e, s1[c]);
          s2[c] = fmaf(b[u], fmaf(e * e, c1[c], -c0[c]), s2[c]);
        }
      }
    }
  }

  // combine the TY row-threads of each column in a fixed order
  __shared__ float red[TY > 1 ? TY : 1][TX][2 * VEC];
  if (TY > 1) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      red[ty][tx][c] = s1[c];
      red[ty][tx][VEC + c] = s2[c];
    }
    __syncthreads();
    if (ty == 0) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        float t1 = red[0][tx][c], t2 = red[0][tx][VEC + c];
        for (int y = 1; y < TY; ++y) {
          t1 += red[y][tx][c];
          t2 += red[y][tx][VEC + c];
        }
        s1[c] = t1;
        s2[c] = t2;
      }
    }
  }
  if (ty == 0 && active) {
    float* p1 = partial + ((int64_t)blockIdx.y * 2 + 0) * D + col;
    float* p2 = partial + ((int64_t)blockIdx.y * 2 + 1) * D + col;
#pragma unroll
    for (int c = 0; c < VEC; ++c) {
      if (col + c < D) {
        p1[c] = s1[c];
        p2[c] = s2[c];
      }
    }
  }
}

__global__ void __launch_bounds__(256) grad_finalize_kernel(const float* __restrict__ partial, int n_chunks, int64_t D, float scale_mu,
                                                            float scale_sigma, float* __restrict__ out_mu, float* __restrict__ out_sigma,
                                                            int64_t item_stride_partial = 0) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= D) return;
  partial += (int64_t)blockIdx.y * item_stride_partial;  // batched: blockIdx.y = item, outputs contiguous [items][D]
  out_mu += (int64_t)blockIdx.y * D;
  out_sigma += (int64_t)blockIdx.y * D;
  float t1 = 0.0f, t2 = 0.0f;
  for (int c = 0; c < n_chunks; ++c) {
    t1 += partial[((int64_t)c * 2 + 0) * D + j];
    t2 += partial[((int64_t)c * 2 + 1) * D + j];
  }
  out_mu[j] = t1 * scale_mu;
  out_sigma[j] = t2 * scale_sigma;
}

// The same finalisation for the sharded generation: this rank's (grad_mu | grad_sigma) goes into slot `rank` of EVERY peer's
// slot array (world x 2D floats) and the last CTA raises this rank's flag on every peer -- the send half of the all-reduce,
// fused into the kernel that produces the data (the receive half is peer_reduce_kernel, evok_peer.cu).
struct GradPush {
  PeerSink sink;
  const unsigned long long* epoch;
  unsigned int* done;
};

__global__ void __launch_bounds__(256) grad_finalize_push_kernel(const float* __restrict__ partial, int n_chunks, int64_t D, float scale_mu,
                                                                 float scale_sigma, const __grid_constant__ PeerSink sink,
                                                                 const unsigned long long* epoch, unsigned int* done) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j < D) {
    float t1 = 0.0f, t2 = 0.0f;
    for (int c = 0; c < n_chunks; ++c) {
      t1 += partial[((int64_t)c * 2 + 0) * D + j];
      t2 += partial[((int64_t)c * 2 + 1) * D + j];
    }
    t1 *= scale_mu;
    t2 *= scale_sigma;
    for (int p = 0; p < sink.world; ++p) {
      float* slot = static_cast<float*>(sink.data[p]) + (int64_t)sink.rank * 2 * D;
      slot[j] = t1;
      slot[D + j] = t2;
    }
  }
  peer_signal_tail(sink, epoch, done);
}

static int launch_finalize(const float* partial, int n_chunks, int64_t D, float scale_mu, float scale_sigma, float* out_mu, float* out_sigma,
                           const GradPush* push, cudaStream_t st) {
  const unsigned grid = (unsigned)((D + 255) / 256);
  if (push) grad_finalize_push_kernel<<<grid, 256, 0, st>>>(partial, n_chunks, D, scale_mu, scale_sigma, push->sink, push->epoch, push->done);
  else grad_finalize_kernel<<<grid, 256, 0, st>>>(partial, n_chunks, D, scale_mu, scale_sigma, out_mu, out_sigma);
  EVOK_CHECK_LAUNCH();
  return 0;
}

// ------------------------------------------------------------------------------------------------------------
// TMA-staged variant of the partial kernel: a producer warp streams row segments global -> shared with 1-D bulk async
// copies (cp.async.bulk, completion counted on an mbarrier), S stages of R rows x 4 KB deep, so each SM keeps
// 2 CTAs x S x R x 4 KB in flight without spending registers or issue slots on loads; 8 consumer warps read the staged
// rows from shared memory (128-bit, conflict free) and accumulate.  Column tile = 1024 columns, one float4 per thread.
// ------------------------------------------------------------------------------------------------------------
#ifndef EVOK_GRAD_TMA_ROWS
#define EVOK_GRAD_TMA_ROWS 4
#endif
#ifndef EVOK_GRAD_TMA_STAGES
#define EVOK_GRAD_TMA_STAGES 4
#endif
#ifndef EVOK_GRAD_TMA_CTAS_PER_SM
#define EVOK_GRAD_TMA_CTAS_PER_SM 3
#endif
constexpr int kTmaRows = EVOK_GRAD_TMA_ROWS;
constexpr int kTmaStages = EVOK_GRAD_TMA_STAGES;
constexpr int kTmaCols = 1024;
constexpr int kTmaConsumers = 256;
constexpr int kTmaThreads = kTmaConsumers + 32;
constexpr size_t kTmaSmemBytes = (size_t)kTmaStages * kTmaRows * kTmaCols * sizeof(float) + 2 * kTmaStages * sizeof(uint64_t) + 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool SYM>
__global__ void __launch_bounds__(kTmaThreads, EVOK_GRAD_TMA_CTAS_PER_SM)
    grad_partial_tma_kernel(int form, const float* __restrict__ X, int64_t ldx, const float* __restrict__ w, const float* __restrict__ mu,
                            const float* __restrict__ sigma, int64_t n_units, int64_t D, int64_t units_per_chunk,
                            float* __restrict__ partial) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tiles = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kTmaStages * kTmaRows * kTmaCols * sizeof(float));
  uint64_t* empty = full + kTmaStages;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t col0 = (int64_t)blockIdx.x * kTmaCols;
  const int64_t width = min((int64_t)kTmaCols, D - col0);  // columns of this tile (multiple of 4)
  const int64_t r_begin = (int64_t)blockIdx.y * units_per_chunk;
  const int64_t r_end = min(n_units, r_begin + units_per_chunk);
  const int64_t n_rows = r_end - r_begin;
  const int64_t n_groups = (n_rows + kTmaRows - 1) / kTmaRows;
  const int64_t row_stride = (SYM ? 2 : 1) * ldx;

  if (tid == 0) {
    for (int s = 0; s < kTmaStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kTmaConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kTmaConsumers / 32) {
    // ===== producer warp: one elected lane issues the bulk copies =====
    if (lane == 0) {
      const uint32_t row_bytes = (uint32_t)(width * sizeof(float));
      for (int64_t g = 0; g < n_groups; ++g) {
        const int s = (int)(g % kTmaStages);
        const uint32_t use = (uint32_t)(g / kTmaStages);
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        const int64_t r0 = r_begin + g * kTmaRows;
        const int rows = (int)min((int64_t)kTmaRows, r_end - r0);
        mbar_expect_tx(&full[s], rows * row_bytes);
        float* dst = tiles + (size_t)s * kTmaRows * kTmaCols;
        for (int i = 0; i < rows; ++i) bulk_load(dst + (size_t)i * kTmaCols, X + (r0 + i) * row_stride + col0, row_bytes, &full[s]);
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int64_t col = col0 + (int64_t)tid * 4;
  const bool active = col < D;
  float m[4], c1[4], c0[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float sg = active ? __ldg(sigma + col + c) : 1.0f;
    m[c] = active ? __ldg(mu + col + c) : 0.0f;
    if (form == EVOK_GRAD_EXP) {
      c1[c] = __fdiv_rn(1.0f, sg * sg);
      c0[c] = 1.0f;
    } else if (form == EVOK_GRAD_MOMENTS) {
      c1[c] = 1.0f;
      c0[c] = 0.0f;
    } else {
      c1[c] = __fdiv_rn(1.0f, sg);
      c0[c] = sg;
    }
  }
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};

  for (int64_t g = 0; g < n_groups; ++g) {
    const int s = (int)(g % kTmaStages);
    const uint32_t use = (uint32_t)(g / kTmaStages);
    const int64_t r0 = r_begin + g * kTmaRows;
    const int rows = (int)min((int64_t)kTmaRows, r_end - r0);
    float a[kTmaRows], b[kTmaRows];
#pragma unroll
    for (int i = 0; i < kTmaRows; ++i) {
      a[i] = b[i] = 0.0f;
      if (i < rows) {
        if (SYM) {
          const float wp = __ldg(w + 2 * (r0 + i)), wm = __ldg(w + 2 * (r0 + i) + 1);
          a[i] = 0.5f * (wp - wm);
          b[i] = 0.5f * (wp + wm);
        } else {
          a[i] = b[i] = __ldg(w + r0 + i);
        }
      }
    }
    mbar_wait(&full[s], use & 1);
    const float4* tile = reinterpret_cast<const float4*>(tiles + (size_t)s * kTmaRows * kTmaCols) + tid;
    if (active) {
#pragma unroll
      for (int i = 0; i < kTmaRows; ++i) {
        if (i < rows) {
          const float4 v = tile[(size_t)i * (kTmaCols / 4)];
          const float x[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float e = x[c] - m[c];
            s1[c] = fmaf(a[i], e, s1[c]);
            s2[c] = fmaf(b[i], fmaf(e * e, c1[c], -c0[c]), s2[c]);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (active) {
    float* p1 = partial + ((int64_t)blockIdx.y * 2 + 0) * D + col;
    float* p2 = partial + ((int64_t)blockIdx.y * 2 + 1) * D + col;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      p1[c] = s1[c];
      p2[c] = s2[c];
    }
  }
}

struct GradPlan {
  int vec, tx, n_coltiles, n_chunks;
  int64_t units_per_chunk;
};

static GradPlan plan_grad(int64_t n_units, int64_t D, bool vec_ok) {
  GradPlan p;
  p.vec = vec_ok ? 4 : 1;
  const int64_t col_threads = (D + p.vec - 1) / p.vec;
  p.tx = 32;
  while (p.tx < kGradThreads && p.tx < col_threads) p.tx <<= 1;
  p.n_coltiles = (int)((col_threads + p.tx - 1) / p.tx);
  const int ty = kGradThreads / p.tx;
  int64_t chunks = (int64_t)kNumSMs * EVOK_GRAD_CTAS_PER_SM / p.n_coltiles;  // one wave of resident CTAs
  // at least 16 unrolled iterations per CTA: fewer, fatter chunks keep the fixed-order finalisation short for small populations
  const int64_t max_useful = (n_units + (int64_t)ty * kGradUnroll * 16 - 1) / ((int64_t)ty * kGradUnroll * 16);
  if (chunks > max_useful) chunks = max_useful;
  if (chunks < 1) chunks = 1;
  if (chunks > 65535) chunks = 65535;
  p.units_per_chunk = (n_units + chunks - 1) / chunks;
  p.n_chunks = (int)((n_units + p.units_per_chunk - 1) / p.units_per_chunk);
  if (p.n_chunks < 1) p.n_chunks = 1;
  return p;
}

template <int VEC, bool SYM, bool REGEN>
static void launch_partial(const GradPlan& p, int form, const float* X, int64_t ldx, const float* w, const float* mu, const float* sigma,
                           int64_t n_units, int64_t D, uint64_t unit0, uint64_t seed, uint64_t stream_id, const uint32_t* stream_off, float* partial,
                           cudaStream_t st, int64_t n_items = 1, const GradItems* items = nullptr) {
  dim3 grid(p.n_coltiles, p.n_chunks, (unsigned)n_items);
  const PhiloxKey key = make_philox_key(seed, stream_id);
  const GradItems it = items ? *items : GradItems{0, 0, 0, 0, 0};
#define EVOK_LAUNCH_TX(TXV)                                                                                                         \
  grad_partial_kernel<VEC, TXV, SYM, REGEN><<<grid, kGradThreads, 0, st>>>(form, X, ldx, w, mu, sigma, n_units, D, p.units_per_chunk, \
                                                                           unit0, key, stream_off, partial, it)
  switch (p.tx) {
    case 32: EVOK_LAUNCH_TX(32); break;
    case 64: EVOK_LAUNCH_TX(64); break;
    case 128: EVOK_LAUNCH_TX(128); break;
    default: EVOK_LAUNCH_TX(256); break;
  }
#undef EVOK_LAUNCH_TX
}

static int grad_impl(int form, const float* X, int64_t ldx, const float* w, const float* mu, const float* sigma, int64_t row0, int64_t n_rows,
                     int64_t D, bool regen, uint64_t seed, uint64_t stream_id, const uint32_t* stream_off, float scale_mu, float scale_sigma, float* out_mu,
                     float* out_sigma, void* ws, size_t ws_bytes, void* stream, const GradPush* push = nullptr) {
  if (!w || !mu || !sigma || !ws || (!regen && !X) || (!push && (!out_mu || !out_sigma))) return EVOK_E_NULLPTR;
  if (form < EVOK_GRAD_SEPARABLE || form > EVOK_GRAD_MOMENTS) return EVOK_E_BADENUM;
  if (n_rows < 0 || D <= 0 || row0 < 0 || (!regen && ldx < D)) return EVOK_E_BADSIZE;
  const bool sym = form == EVOK_GRAD_SYMMETRIC;
  if (sym && ((n_rows & 1) || (row0 & 1))) return EVOK_E_ODDROWS;
  if (ws_bytes < evok_grad_workspace_bytes(n_rows, D)) return EVOK_E_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_units = sym ? n_rows / 2 : n_rows;
  const bool vec_ok = regen ? true : ((D % 4 == 0) && (ldx % 4 == 0) && aligned16(X));
  // symmetric sampling keys its counters by direction; the regenerating kernel must use the same unit index
  const uint64_t unit0 = (uint64_t)(sym ? row0 / 2 : row0);
  GradPlan p = plan_grad(n_units, D, vec_ok);
  float* partial = (float*)ws;
  if (n_units == 0 && push) return launch_finalize(partial, 0, D, scale_mu, scale_sigma, nullptr, nullptr, push, st);  // zeros + this rank's flag
  if (n_units == 0) {
    cudaMemsetAsync(out_mu, 0, (size_t)D * 4, st);
    cudaMemsetAsync(out_sigma, 0, (size_t)D * 4, st);
    return 0;
  }
  // read per call (a getenv is ~100 ns): the parity tests run both implementations in one process
  const char* tma_env = getenv("EVOK_GRAD_TMA");
  const int use_tma = tma_env ? atoi(tma_env) : EVOK_GRAD_TMA_DEFAULT;
  if (use_tma && !regen && vec_ok && form != EVOK_GRAD_MOMENTS && D >= 512 && n_units >= 4096) {
    const int n_coltiles = (int)((D + kTmaCols - 1) / kTmaCols);
    int64_t chunks = (int64_t)kNumSMs * EVOK_GRAD_TMA_CTAS_PER_SM / n_coltiles;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    const int64_t upc = (n_units + chunks - 1) / chunks;
    const int n_chunks = (int)((n_units + upc - 1) / upc);
    dim3 grid(n_coltiles, n_chunks);
    if (sym) {
      cudaFuncSetAttribute(grad_partial_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes);
      grad_partial_tma_kernel<true><<<grid, kTmaThreads, kTmaSmemBytes, st>>>(form, X, ldx, w, mu, sigma, n_units, D, upc, partial);
    } else {
      cudaFuncSetAttribute(grad_partial_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTmaSmemBytes);
      grad_partial_tma_kernel<false><<<grid, kTmaThreads, kTmaSmemBytes, st>>>(form, X, ldx, w, mu, sigma, n_units, D, upc, partial);
    }
    EVOK_CHECK_LAUNCH();
    return launch_finalize(partial, n_chunks, D, scale_mu, scale_sigma, out_mu, out_sigma, push, st);
  }
  if (regen) {
    if (sym) launch_partial<4, true, true>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
    else launch_partial<4, false, true>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
  } else if (vec_ok) {
    if (sym) launch_partial<4, true, false>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
    else launch_partial<4, false, false>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
  } else {
    if (sym) launch_partial<1, true, false>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
    else launch_partial<1, false, false>(p, form, X, ldx, w, mu, sigma, n_units, D, unit0, seed, stream_id, stream_off, partial, st);
  }
  EVOK_CHECK_LAUNCH();
  return launch_finalize(partial, p.n_chunks, D, scale_mu, scale_sigma, out_mu, out_sigma, push, st);
}

}  // namespace evok

using namespace evok;

extern "C" EVOK_API size_t evok_grad_workspace_bytes(int64_t n_rows, int64_t D) {
  (void)n_rows;
  if (D <= 0) return 256;
  // n_chunks * n_coltiles <= kMaxResidentCtas/2 + n_coltiles and every column tile spans <= 1024 columns
  return ((size_t)kMaxResidentCtas * 1024 + 2 * (size_t)D + 64) * sizeof(float);
}

extern "C" EVOK_API int evok_grad(int form, const float* X, int64_t ldx, const float* w, const float* mu, const float* sigma, int64_t n_rows,
                         int64_t D, float scale_mu, float scale_sigma, float* out_mu, float* out_sigma, void* ws, size_t ws_bytes,
                         void* stream) {
  return grad_impl(form, X, ldx, w, mu, sigma, 0, n_rows, D, false, 0, 0, nullptr, scale_mu, scale_sigma, out_mu, out_sigma, ws, ws_bytes, stream);
}

extern "C" EVOK_API int evok_grad_regen(int form, const float* w, const float* mu, const float* sigma, int64_t row0, int64_t n_rows, int64_t D,
                               uint64_t seed, uint64_t stream_id, const uint32_t* stream_offset_dev, float scale_mu, float scale_sigma,
                               float* out_mu, float* out_sigma, void* ws, size_t ws_bytes, void* stream) {
  return grad_impl(form, nullptr, 0, w, mu, sigma, row0, n_rows, D, true, seed, stream_id, stream_offset_dev, scale_mu, scale_sigma, out_mu, out_sigma, ws,
                   ws_bytes, stream);
}

extern "C" EVOK_API int evok_grad_push(int form, const float* X, int64_t ldx, const float* w, const float* mu, const float* sigma, int64_t row0,
                                       int64_t n_rows, int64_t D, uint64_t seed, uint64_t stream_id, const uint32_t* stream_offset_dev,
                                       float scale_mu, float scale_sigma, int world, int rank, void* const* peer_slots, void* const* peer_flags,
                                       const uint64_t* epoch_dev, uint32_t* done_dev, void* ws, size_t ws_bytes, void* stream) {
  if (!peer_slots || !peer_flags || !epoch_dev || !done_dev) return EVOK_E_NULLPTR;
  if (world < 1 || world > EVOK_MAX_PEERS || rank < 0 || rank >= world) return EVOK_E_BADSIZE;
  GradPush push{};
  push.sink.world = world;
  push.sink.rank = rank;
  for (int p = 0; p < world; ++p) {
    if (!peer_slots[p] || !peer_flags[p]) return EVOK_E_NULLPTR;
    push.sink.data[p] = peer_slots[p];
    push.sink.flags[p] = static_cast<unsigned long long*>(peer_flags[p]);
  }
  push.epoch = reinterpret_cast<const unsigned long long*>(epoch_dev);
  push.done = done_dev;
  return grad_impl(form, X, X ? ldx : 0, w, mu, sigma, row0, n_rows, D, X == nullptr, seed, stream_id, stream_offset_dev, scale_mu, scale_sigma, nullptr,
                   nullptr, ws, ws_bytes, stream, &push);
}

// Batched searches: n_items independent weighted column reductions in ONE launch chain (blockIdx.z = item).  X: [items][n_rows][D]
// (item stride item_stride_x elements), w: [items][n_rows] contiguous, mu / sigma: item strides (0 = shared by all items),
// outputs contiguous [items][D].  Same arithmetic as evok_grad per item (LDG kernel, fixed-order two-stage reduction).
extern "C" EVOK_API size_t evok_grad_batched_workspace_bytes(int64_t n_items, int64_t n_rows, int64_t D) {
  if (n_items <= 0 || D <= 0) return 256;
  (void)n_rows;
  // all items together use about one wave of CTAs: sum over items of n_chunks * 2 * D <= (592 / coltiles) * 2 * D + (2 per item of slack) * 2 * D
  return ((size_t)kMaxResidentCtas * 1024 + 4 * (size_t)n_items * (size_t)D + 64) * sizeof(float);
}

extern "C" EVOK_API int evok_grad_batched(int form, const float* X, int64_t item_stride_x, int64_t ldx, const float* w, const float* mu,
                                          int64_t item_stride_mu, const float* sigma, int64_t item_stride_sigma, int64_t n_items, int64_t n_rows,
                                          int64_t D, float scale_mu, float scale_sigma, float* out_mu, float* out_sigma, void* ws, size_t ws_bytes,
                                          void* stream) {
  if (!X || !w || !mu || !sigma || !out_mu || !out_sigma || !ws) return EVOK_E_NULLPTR;
  if (form < EVOK_GRAD_SEPARABLE || form > EVOK_GRAD_MOMENTS) return EVOK_E_BADENUM;
  if (n_items < 0 || n_items > 65535 || n_rows < 0 || D <= 0 || ldx < D) return EVOK_E_BADSIZE;
  const bool sym = form == EVOK_GRAD_SYMMETRIC;
  if (sym && (n_rows & 1)) return EVOK_E_ODDROWS;
  if (n_items == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_units = sym ? n_rows / 2 : n_rows;
  if (n_units == 0) {
    cudaMemsetAsync(out_mu, 0, (size_t)n_items * D * 4, st);
    cudaMemsetAsync(out_sigma, 0, (size_t)n_items * D * 4, st);
    return 0;
  }
  const bool vec_ok = (D % 4 == 0) && (ldx % 4 == 0) && aligned16(X) && item_stride_x % 4 == 0 && item_stride_mu % 4 == 0 && item_stride_sigma % 4 == 0;
  GradPlan p = plan_grad(n_units, D, vec_ok);
  // the batch fills the GPU: fewer row chunks per item keep the fixed-order finalisation short
  int64_t chunks = ((int64_t)kNumSMs * EVOK_GRAD_CTAS_PER_SM + (int64_t)p.n_coltiles * n_items - 1) / ((int64_t)p.n_coltiles * n_items);
  if (chunks < p.n_chunks) {
    if (chunks < 1) chunks = 1;
    p.units_per_chunk = (n_units + chunks - 1) / chunks;
    p.n_chunks = (int)((n_units + p.units_per_chunk - 1) / p.units_per_chunk);
  }
  GradItems items;
  items.x = item_stride_x;
  items.w = n_rows;
  items.mu = item_stride_mu;
  items.sigma = item_stride_sigma;
  items.partial = (int64_t)p.n_chunks * 2 * D;
  if (ws_bytes < (size_t)n_items * items.partial * sizeof(float)) return EVOK_E_WORKSPACE;
  float* partial = (float*)ws;
  if (vec_ok) {
    if (sym) launch_partial<4, true, false>(p, form, X, ldx, w, mu, sigma, n_units, D, 0, 0, 0, nullptr, partial, st, n_items, &items);
    else launch_partial<4, false, false>(p, form, X, ldx, w, mu, sigma, n_units, D, 0, 0, 0, nullptr, partial, st, n_items, &items);
  } else {
    if (sym) launch_partial<1, true, false>(p, form, X, ldx, w, mu, sigma, n_units, D, 0, 0, 0, nullptr, partial, st, n_items, &items);
    else launch_partial<1, false, false>(p, form, X, ldx, w, mu, sigma, n_units, D, 0, 0, 0, nullptr, partial, st, n_items, &items);
  }
  EVOK_CHECK_LAUNCH();
  grad_finalize_kernel<<<dim3((unsigned)((D + 255) / 256), (unsigned)n_items), 256, 0, st>>>(partial, p.n_chunks, D, scale_mu, scale_sigma, out_mu,
                                                                                             out_sigma, items.partial);
  EVOK_CHECK_LAUNCH();
  return 0;
}
