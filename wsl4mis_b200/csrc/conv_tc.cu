// tcgen05 implicit-GEMM convolution for sm_100a (3x3 pad 1 / 1x1) on channels-last bf16 activations.
//
//   D[128 pixels, NT couts] (fp32, TMEM) = sum_{tap} sum_{kblock} A_tap[128 x KBLK] * B_tap[KBLK x NT]
//
//   * A: one TMA 4-D box (KBLK channels x 16 w x 8 h x 1 image) per (tap, k-block) at the tap-shifted pixel
//     coordinate; TMA's out-of-bounds zero fill IS the convolution's zero padding (im2col-free).
//   * B: packed weights [tap][Cout][Cin] bf16 (K-major), one 3-D box (KBLK x NT x 1) per (tap, k-block).
//   * both land in 128B/64B/32B-swizzled shared memory (swizzle span = KBLK*2 bytes) and are consumed by
//     tcgen05.mma.cta_group::1.kind::f16 (M=128, N=NT, K=16) issued by one elected thread; fp32 accumulators in TMEM.
//   * warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..5 = epilogue
//     (tcgen05.ld 32x32b -> +bias -> bf16 NHWC / fp32 NCHW stores).
//   * mbarrier ring (full/empty) of STAGES stages between TMA and MMA; tcgen05.commit releases a stage.
//
// The same kernel is the data-gradient (dgrad) kernel when fed tap-flipped, transposed weights, and the 1x1
// convolution with TAPS = 1.  Two source tensors model torch.cat([skip, up], 1) without materialising it.
// Reference call sites: networks/unet.py:19,23 (3x3), :55 (1x1), :67 (concat), :120 (out_conv).
#include "common.cuh"
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <string.h>
#include <stdlib.h>

namespace {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must abort the kernel (trap -> launch error), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    if (spin > (1u << 24)) {
      printf("wsl conv_tc: mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Issue-path helpers.  The MMA warp stays converged; ONE elected lane issues, and every operand of the issue loop is a
// warp-uniform value (read through lane 0), so the compiler keeps descriptors in uniform registers and emits back-to-back
// UTCHMMA instead of a per-instruction ELECT / R2UR / branch sequence (measured on the B200: ~80 -> ~50 cycles per
// M=128,N=16 MMA; profiles/micro/mma_rate.cu).  A descriptor is {lo = addr>>4 | LBO<<16, hi = constant}.
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t e;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
  return e;
}
__device__ __forceinline__ uint32_t uniform(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ void umma_f16_w(uint32_t tmem_d, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// high word of a shared-memory descriptor: SBO | version 1 | swizzle layout
template <int SW>
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) {
  return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14) | (((SW == 128) ? 2u : (SW == 64) ? 4u : 6u) << 29);
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major operand, swizzle span = SW bytes (32/64/128):
//   canonical layout ((8,m),(T,2)):((SW/16,SBO),(1,.)) in 16-byte units -> 8-row groups SBO = 8*SW bytes apart.
template <int SW>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t saddr) {
  constexpr uint64_t layout = (SW == 128) ? 2 : (SW == 64) ? 4 : 6;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);          // start address  [0,14)
  d |= (uint64_t)1 << 16;                            // LBO (ignored for swizzled K-major) [16,30)
  d |= (uint64_t)((8 * SW) >> 4) << 32;              // SBO [32,46)
  d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
  d |= layout << 61;                                 // swizzle mode
  return d;
}

// instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=n
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int a_mn_major = 0, int b_mn_major = 0, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// the same descriptor for the run-time 16-bit operand type: dt 0 = bf16 (format code 1), dt 2 = fp16 (format code 0);
// kind::f16 runs both at the same rate, fp16 carries an 11-bit instead of an 8-bit mantissa
__device__ __forceinline__ uint32_t idesc_for(uint32_t idesc_bf16, int dt) {
  return dt == 2 ? (idesc_bf16 & ~((1u << 7) | (1u << 10))) : idesc_bf16;
}

struct ConvTcParams {
  int N, H, W;
  int C0, C1;        // channels of source 0 / source 1 (C1 may be 0)
  int CoutP;         // padded output channels (rows of the packed weight)
  int CoutStore;     // channels stored per pixel (bf16 NHWC) or real channels (fp32 NCHW)
  int taps, ks;
  int tiles_x, tiles_y;
  int out_mode;       // 0: 16-bit NHWC, 1: fp32 NCHW, 2: fp32 NHWC
  int dt;             // 16-bit operand / storage type: 0 bf16, 2 fp16
  int dil;            // dilation (PNet2D's blocks, networks/pnet.py:25-28): tap (dy, dx) reads pixel (y + dil*dy, x + dil*dx)
  const float* bias;
  const float* out_scale;   // optional device scalar multiplied into the accumulator before the bias (split mode: 2^-k of the staged operand)
  void* out;
};

constexpr int TILE_W = 16, TILE_H = 8, TILE_M = 128;
constexpr int NUM_THREADS = 192;

template <int KBLK, int NT, int STAGES>
struct ConvSmem {
  static constexpr int A_BYTES = TILE_M * KBLK * 2;
  static constexpr int B_BYTES = NT * KBLK * 2;
  static constexpr int STAGE_BYTES = ((A_BYTES + B_BYTES + 1023) / 1024) * 1024;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int KBLK, int NT, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0,
                                                                 const __grid_constant__ CUtensorMap map_a1,
                                                                 const __grid_constant__ CUtensorMap map_b, const ConvTcParams p) {
  using S = ConvSmem<KBLK, NT, STAGES>;
  constexpr int SW = KBLK * 2;
  constexpr uint32_t TMEM_COLS = NT < 32 ? 32 : NT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x;
  const int n = tile / (p.tiles_x * p.tiles_y);
  const int tr = tile - n * p.tiles_x * p.tiles_y;
  const int y0 = (tr / p.tiles_x) * TILE_H, x0 = (tr % p.tiles_x) * TILE_W;
  const int n0 = blockIdx.y * NT;
  const int pad = p.ks >> 1;
  const int kb0 = p.C0 / KBLK, kb1 = p.C1 / KBLK;
  const int iters = p.taps * (kb0 + kb1);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_b);
    if (kb1) prefetch_tmap(&map_a1);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < p.taps; ++t) {
        const int dy = (t / p.ks - pad) * p.dil, dx = (t % p.ks - pad) * p.dil;
        for (int kb = 0; kb < kb0 + kb1; ++kb, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * S::STAGE_BYTES;
          uint8_t* b_dst = a_dst + S::A_BYTES;
          mbar_expect_tx(&full_bar[s], S::A_BYTES + S::B_BYTES);
          if (kb < kb0) tma_load_4d(&map_a0, &full_bar[s], a_dst, kb * KBLK, x0 + dx, y0 + dy, n);
          else          tma_load_4d(&map_a1, &full_bar[s], a_dst, (kb - kb0) * KBLK, x0 + dx, y0 + dy, n);
          tma_load_3d(&map_b, &full_bar[s], b_dst, kb * KBLK, n0, t);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = idesc_for(make_idesc_bf16(NT), p.dt);
    for (int it = 0; it < iters; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        const uint64_t adesc = make_kmajor_desc<SW>(a_addr);
        const uint64_t bdesc = make_kmajor_desc<SW>(b_addr);
#pragma unroll
        for (int k = 0; k < KBLK / 16; ++k) {
          // advance 16 elements (32 bytes) along K inside the swizzle span: +2 in the 16-byte start-address field
          umma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);                 // frees the smem stage when these MMAs retire
        if (it == iters - 1) umma_commit(accum_bar); // accumulator complete -> epilogue
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int m = q * 32 + lane;            // accumulator row = pixel index inside the tile
    const int gy = y0 + m / TILE_W, gx = x0 + m % TILE_W;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const bool inb = (gy < p.H) && (gx < p.W);
#pragma unroll 1
    for (int c = 0; c < NT; c += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      if (p.out_scale) {
        const float sc = *p.out_scale;
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] *= sc;
      }
      if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += p.bias[n0 + c + j];
      }
      if (!inb) continue;
      if (p.out_mode == 0) {
        if (n0 + c < p.CoutStore) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (((long long)n * p.H + gy) * p.W + gx) * p.CoutStore + n0 + c;
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
          reinterpret_cast<uint4*>(o)[0] = pack8_dt(lo, p.dt);
          if (n0 + c + 8 < p.CoutStore) reinterpret_cast<uint4*>(o)[1] = pack8_dt(hi, p.dt);
        }
      } else if (p.out_mode == 2) {
        if (n0 + c < p.CoutStore) {
          float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + (((long long)n * p.H + gy) * p.W + gx) * p.CoutStore + n0 + c);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n0 + c + 4 * j < p.CoutStore) o[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      } else {
        float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (n0 + c + j < p.CoutStore) o[(((long long)n * p.CoutStore + n0 + c + j) * p.H + gy) * p.W + gx] = v[j];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------
// host: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(f);
  });
  return fn;
}

CUtensorMapSwizzle swizzle_for(int kblk) {
  return kblk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : kblk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

struct MapKey {
  const void* ptr; long long d[4]; int b[4]; int rank; int sw; int dt; int pitch;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    const uint64_t* w = reinterpret_cast<const uint64_t*>(&k);
    uint64_t h = 1469598103934665603ULL;
    for (size_t i = 0; i < sizeof(MapKey) / 8; ++i) { h ^= w[i]; h *= 1099511628211ULL; }
    return (size_t)h;
  }
};
std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 16-bit tensor (dt 0 = bf16, 2 = fp16), dims innermost-first; strides derived (dense) unless `pitch` (elements between
// consecutive indices of dimension 1, i.e. the channel pitch of a channels-last tensor) is given: a channel-range view of a
// wider tensor.  Cached by (ptr, dims, box, swizzle, type, pitch).
int get_map(const void* ptr, int rank, const long long* dims, const int* box, int kblk, CUtensorMap* out, int dt = 0, int pitch = 0) {
  MapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.rank = rank; key.sw = kblk; key.dt = dt; key.pitch = pitch;
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
  {
    std::lock_guard<std::mutex> g(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) { wsl_set_error("cuTensorMapEncodeTiled unavailable (driver too old?)"); return -3; }
  cuuint64_t gdim[4], gstride[3];
  cuuint32_t bdim[4], estr[4];
  unsigned long long stride = 2;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = (cuuint64_t)dims[i];
    bdim[i] = (cuuint32_t)box[i];
    estr[i] = 1;
    stride *= (unsigned long long)((i == 0 && pitch > 0) ? pitch : dims[i]);
    if (i < rank - 1) gstride[i] = stride;
  }
  CUresult r = enc(out, dt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride, bdim, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(kblk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { wsl_set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -4; }
  std::lock_guard<std::mutex> g(g_map_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

template <int KBLK, int NT, int STAGES>
int launch_conv(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvTcParams& p, cudaStream_t stream) {
  using S = ConvSmem<KBLK, NT, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<KBLK, NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { wsl_set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -5; }
    attr = true;
  }
  dim3 grid(p.N * p.tiles_x * p.tiles_y, p.CoutP / NT);
  conv_tc_kernel<KBLK, NT, STAGES><<<grid, NUM_THREADS, S::TOTAL, stream>>>(a0, a1, b, p);
  return wsl_check_launch("conv_tc");
}

template <int KBLK>
int dispatch_nt(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvTcParams& p, int nt, cudaStream_t stream) {
  switch (nt) {
    case 16:  return launch_conv<KBLK, 16, 6>(a0, a1, b, p, stream);
    case 32:  return launch_conv<KBLK, 32, 6>(a0, a1, b, p, stream);
    case 64:  return launch_conv<KBLK, 64, 6>(a0, a1, b, p, stream);
    case 128: return launch_conv<KBLK, 128, 5>(a0, a1, b, p, stream);
  }
  wsl_set_error("conv_tc: unsupported N tile %d", nt);
  return -1;
}


// ================================================================================================
// conv_tc v2: persistent, weights resident in shared memory, ONE halo load per (tile, k-block)
//
//   * pixel tile = 8 w x (16*MT) h  (MT accumulators of M = 128 rows, row m = h*8 + w);
//   * A: one TMA box (KBLK channels x 10 w x (16*MT+2) h) per k-block -- the 3x3 halo.  The nine taps are NOT
//     reloaded: tap (dy,dx) of sub-tile j is the same shared-memory tile viewed through a UMMA descriptor whose
//     start address is shifted by ((16*j+dy)*10 + dx) rows and whose 8-row-group stride (SBO) is one halo row
//     (10 pixels).  The 128B/64B/32B swizzle is a function of the absolute shared-memory address (TMA writes and
//     tcgen05 reads agree), so a row-shifted start keeps the pattern consistent.
//   * B: the CTA's whole weight slice [taps][k-blocks][NT x KBLK] is loaded once (it fits: NT is chosen on the
//     host so that taps*Cin*NT*2 <= ~148 KB) and stays resident while the CTA walks its pixel tiles.
//   * TMEM: two sets of MT*NT fp32 columns; the epilogue of tile i overlaps the MMAs of tile i+1.
// L2->SM traffic per output pixel drops from 9*(Cin + NT)*2 B (v1) to ~1.4*Cin*2 B.
// ================================================================================================
template <int SW>
__device__ __forceinline__ uint64_t make_kmajor_desc_sbo(uint32_t saddr, uint32_t sbo_bytes, int base_offset_mode) {
  constexpr uint64_t layout = (SW == 128) ? 2 : (SW == 64) ? 4 : 6;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  if (base_offset_mode) d |= (uint64_t)((saddr >> 7) & 7) << 49;
  d |= layout << 61;
  return d;
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct ConvV2Params {
  int N, H, W;
  int C0, C1;
  int CoutP, CoutStore;
  int tiles_x, tiles_y, ntiles;
  int out_mode;
  int desc_mode;
  int dt;                 // 16-bit operand / storage type: 0 bf16, 2 fp16
  const float* bias;
  void* out;
  float* stat_partials;   // optional [gridDim.x][2][CoutP]: per-CTA sum / sum of squares of the stored (bf16-rounded) outputs
};

// Sum over the 32 lanes of a warp for 16 per-lane values with a transpose-reduction (16 shuffles instead of 80):
// returns, on every lane, the total of value index ((lane >> 1) & 15).
__device__ __forceinline__ float warp_transpose_sum16(const float (&x)[16], int lane) {
  float y[8], z[4], w[2];
  const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = h16 ? x[i] : x[i + 8], keep = h16 ? x[i + 8] : x[i];
    y[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = h8 ? y[i] : y[i + 4], keep = h8 ? y[i + 4] : y[i];
    z[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = h4 ? z[i] : z[i + 2], keep = h4 ? z[i + 2] : z[i];
    w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = h2 ? w[0] : w[1], keep = h2 ? w[1] : w[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

template <int KS, int KBLK, int NT, int MT>
struct ConvV2Cfg {
  static constexpr int PAD = KS / 2;
  static constexpr int TW = 8, TH = 16 * MT;
  static constexpr int HW_ = TW + 2 * PAD, HH_ = TH + 2 * PAD;
  static constexpr int ROWB = KBLK * 2;
  static constexpr int HALO_BYTES = HW_ * HH_ * ROWB;
  static constexpr int A_STAGE = ((HALO_BYTES + 1023) / 1024) * 1024;
  static constexpr int W_SUB = NT * ROWB;                  // one (tap, k-block) weight tile
  static constexpr uint32_t TMEM_COLS = (2 * MT * NT <= 32) ? 32 : (2 * MT * NT <= 64) ? 64 : (2 * MT * NT <= 128) ? 128
                                        : (2 * MT * NT <= 256) ? 256 : 512;
};

// conv_tc2 runs EIGHT epilogue warps (two per TMEM lane quarter, each taking half of the accumulator columns): the
// epilogue (tcgen05.ld, bias, bf16 pack, stores, BatchNorm partial sums) was the pacing stage with four.
template <int NT, int MT>
struct Conv2Epi { static constexpr int WARPS = (NT >= 64 && MT == 1) ? 8 : 4; static constexpr int THREADS = 64 + 32 * WARPS; };

template <int KS, int KBLK, int NT, int MT, int STAGES>
__global__ void __launch_bounds__(Conv2Epi<NT, MT>::THREADS) conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a0,
                                                               const __grid_constant__ CUtensorMap map_a1,
                                                               const __grid_constant__ CUtensorMap map_b, const ConvV2Params p,
                                                               const int w_bytes) {
  using Cfg = ConvV2Cfg<KS, KBLK, NT, MT>;
  constexpr int SW = KBLK * 2, T = KS * KS;
  static_assert(2 * MT * NT <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* w_smem = smem;
  uint8_t* a_smem = smem + w_bytes;                         // w_bytes is a multiple of 1024
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + STAGES * Cfg::A_STAGE);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + STAGES;
  uint64_t* acc_full = bars + 2 * STAGES;
  uint64_t* acc_empty = bars + 2 * STAGES + 2;
  uint64_t* w_bar = bars + 2 * STAGES + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * NT;
  const int kb0 = p.C0 / KBLK, nkb = kb0 + p.C1 / KBLK;
  const int my_tiles = (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], Conv2Epi<NT, MT>::WARPS); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_b);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)(T * nkb * Cfg::W_SUB));
      for (int t = 0; t < T; ++t)
        for (int kb = 0; kb < nkb; ++kb) tma_load_3d(&map_b, w_bar, w_smem + (t * nkb + kb) * Cfg::W_SUB, kb * KBLK, n0, t);
      int it = 0;
      for (int i = 0; i < my_tiles; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int n = tile / (p.tiles_x * p.tiles_y);
        const int tr = tile - n * p.tiles_x * p.tiles_y;
        const int y0 = (tr / p.tiles_x) * Cfg::TH, x0 = (tr % p.tiles_x) * Cfg::TW;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&a_empty[s], ((it / STAGES) & 1) ^ 1);
          mbar_expect_tx(&a_full[s], Cfg::HALO_BYTES);
          if (kb < kb0) tma_load_4d(&map_a0, &a_full[s], a_smem + s * Cfg::A_STAGE, kb * KBLK, x0 - Cfg::PAD, y0 - Cfg::PAD, n);
          else          tma_load_4d(&map_a1, &a_full[s], a_smem + s * Cfg::A_STAGE, (kb - kb0) * KBLK, x0 - Cfg::PAD, y0 - Cfg::PAD, n);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_for(make_idesc_bf16(NT), p.dt);
    constexpr uint32_t a_hi = desc_hi<SW>(Cfg::HW_ * Cfg::ROWB), b_hi = desc_hi<SW>(8 * SW);
    const uint32_t elected = elect_one();
    const uint32_t tmem_u = uniform(tmem_base);
    const uint32_t a_lo0 = (uniform(smem_u32(a_smem)) >> 4) | 0x10000u;      // LBO field = 1 (ignored for swizzled K-major)
    const uint32_t w_lo0 = (uniform(smem_u32(w_smem)) >> 4) | 0x10000u;
    const uint32_t w_tap = (uint32_t)(nkb * Cfg::W_SUB) >> 4;               // descriptor units between two taps of the weights
    mbar_wait(w_bar, 0);
    int it = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int acc = i & 1;
      mbar_wait(&acc_empty[acc], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&a_full[s], (it / STAGES) & 1);
        tc_fence_after();
        if (elected) {
          const uint32_t a_lo = a_lo0 + (uint32_t)(s * (Cfg::A_STAGE >> 4));
          const uint32_t w_lo = w_lo0 + (uint32_t)(kb * (Cfg::W_SUB >> 4));
          const uint32_t d0 = tmem_u + (uint32_t)(acc * MT * NT);
#pragma unroll
          for (int j = 0; j < MT; ++j) {
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const int dy = t / KS, dx = t % KS;
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k)
                umma_f16_w(d0 + (uint32_t)(j * NT), a_lo + (uint32_t)(((((16 * j + dy) * Cfg::HW_ + dx) * Cfg::ROWB) >> 4) + 2 * k), a_hi,
                           w_lo + (uint32_t)t * w_tap + (uint32_t)(2 * k), b_hi, idesc, (kb > 0 || t > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&a_empty[s]);
          if (kb == nkb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    constexpr int NCH = NT / 16;                           // 16-column chunks of the accumulator
    constexpr int EW = Conv2Epi<NT, MT>::WARPS;
    const int half = (warp - 2) >> 2;                      // which half of the chunks this warp owns (EW == 8)
    const int ch_lo = (EW == 8) ? half * (NCH / 2) : 0;
    const int ch_hi = (EW == 8) ? (half + 1) * (NCH / 2) : NCH;
    float rs[NT / 16], rq[NT / 16];       // running per-channel sum / sum of squares (channel = chunk*16 + ((lane>>1)&15))
#pragma unroll
    for (int c = 0; c < NT / 16; ++c) rs[c] = rq[c] = 0.f;
    for (int i = 0; i < my_tiles; ++i) {
      const int tile = blockIdx.x + i * gridDim.x;
      const int n = tile / (p.tiles_x * p.tiles_y);
      const int tr = tile - n * p.tiles_x * p.tiles_y;
      const int y0 = (tr / p.tiles_x) * Cfg::TH, x0 = (tr % p.tiles_x) * Cfg::TW;
      const int acc = i & 1;
      mbar_wait(&acc_full[acc], (i >> 1) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int gy = y0 + 16 * j + m / 8, gx = x0 + (m & 7);
#pragma unroll
        for (int c = 0; c < NT; c += 16) {
          if (c / 16 < ch_lo || c / 16 >= ch_hi) continue;   // warp-uniform: the other epilogue warp of this quarter owns it
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((acc * MT + j) * NT + c), v);
          if (p.bias) {
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) v[jj] += p.bias[n0 + c + jj];
          }
          if (p.out_mode == 0) {
            float lo[8], hi[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) { lo[jj] = v[jj]; hi[jj] = v[8 + jj]; }
            const uint4 plo = pack8_dt(lo, p.dt), phi = pack8_dt(hi, p.dt);
            if (n0 + c < p.CoutStore) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (((long long)n * p.H + gy) * p.W + gx) * p.CoutStore + n0 + c;
              reinterpret_cast<uint4*>(o)[0] = plo;
              if (n0 + c + 8 < p.CoutStore) reinterpret_cast<uint4*>(o)[1] = phi;
            }
            if (p.stat_partials) {   // BatchNorm statistics of the values exactly as stored (rounded to the 16-bit type)
              float r[16], r2[16];
              unpack8_dt(plo, lo, p.dt);
              unpack8_dt(phi, hi, p.dt);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) { r[jj] = lo[jj]; r[8 + jj] = hi[jj]; }
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) r2[jj] = r[jj] * r[jj];
              rs[c / 16] += warp_transpose_sum16(r, lane);
              rq[c / 16] += warp_transpose_sum16(r2, lane);
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (n0 + c + jj < p.CoutStore) o[(((long long)n * p.CoutStore + n0 + c + jj) * p.H + gy) * p.W + gx] = v[jj];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
    if (p.stat_partials) {
      // combine the four epilogue warps (fixed order) and write this CTA's partial row
      __shared__ float s_stat[4][2][NT];
      if ((lane & 1) == 0) {
#pragma unroll
        for (int c = 0; c < NT / 16; ++c) {
          if (c < ch_lo || c >= ch_hi) continue;
          s_stat[q][0][c * 16 + ((lane >> 1) & 15)] = rs[c];
          s_stat[q][1][c * 16 + ((lane >> 1) & 15)] = rq[c];
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      const int t = threadIdx.x - 64;      // index over the epilogue warps
      for (int c = t; c < 2 * NT; c += 32 * EW) {
        const int which = c / NT, ch = c % NT;
        const float a = ((s_stat[0][which][ch] + s_stat[1][which][ch]) + s_stat[2][which][ch]) + s_stat[3][which][ch];
        p.stat_partials[((size_t)blockIdx.x * 2 + which) * p.CoutP + n0 + ch] = a;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ================================================================================================
// tcgen05 weight gradient:  dW[co][ci][t] += sum_{pixels} dY[p][co] * X[p + tap_t][ci]
//
//   D_t[M = 128 couts, N = CWB cins] (fp32, TMEM) += A^T[K = 128 pixels x M] * B_t[K = 128 pixels x N]
//
// Both operands are "MN-major" for the tensor core: K (pixels) indexes shared-memory rows, the channel (M or N)
// is contiguous inside a row -- exactly what a channels-last TMA box (channels x 16 w x 8 h) produces.  A (dY) is
// loaded once per 128-pixel chunk, B (the tap-shifted X box, TMA zero fill = padding) once per tap of the CTA's
// tap group; each CTA owns (Cout tile of 128) x (Cin block of CWB) x (TG taps), loops over its share of the
// N*H*W/128 pixel chunks (split-K), keeps TG accumulators in TMEM and finally adds them into the fp32
// torch-layout gradient with red.global.add.f32.
//   CWA/CWB = channels per A/B box (16/32/64 -> swizzle 32/64/128 B), NA = A boxes per stage (1 or 2).
// When Cout < 128 the MMA still runs M = 128 (same tensor-pipe cost as M = 64): the missing row groups alias
// whatever shared memory follows the A box; those accumulator rows are never read.
// ================================================================================================
template <int SW>
__device__ __forceinline__ uint64_t make_mnmajor_desc(uint32_t saddr, uint32_t lbo_bytes) {
  constexpr uint64_t layout = (SW == 128) ? 2 : (SW == 64) ? 4 : 6;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;   // stride between SW-byte wide channel groups
  d |= (uint64_t)((8 * SW) >> 4) << 32;                // stride between 8-pixel row groups
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

struct WgradTcParams {
  int N, H, W;
  int C0, C1;          // source channels
  int CoutP, CoutReal;
  int taps, ks;
  int tiles_x, tiles_y, nchunks;
  int n_tiles0, n_tiles1;   // Cin blocks in source 0 / 1
  int tap_groups;
  int dt;               // 16-bit operand type: 0 bf16, 2 fp16
  const float* scale_a;  // optional device scalars multiplied into the result (split mode: 2^-k of the staged dY / X)
  const float* scale_b;
  float* dw;
  float* partials;       // optional [gridDim.y][gridDim.x][TG*CWB][128]: per-CTA partial tiles for the fixed-order finalize
  int dil;               // dilation of the convolution
};

template <int CWA, int NA, int CWB, int TG, int STAGES>
struct WgradSmem {
  static constexpr int A_BOX = TILE_M * CWA * 2;
  static constexpr int B_BOX = TILE_M * CWB * 2;
  static constexpr int STAGE_BYTES = ((NA * A_BOX + TG * B_BOX + 1023) / 1024) * 1024;
  static constexpr int SLACK = (128 / CWA) * A_BOX;   // aliased row groups of the last stage stay inside the allocation
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SLACK + 1024 + 256;
};

template <int CWA, int NA, int CWB, int TG, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                  const __grid_constant__ CUtensorMap map_x0,
                                                                  const __grid_constant__ CUtensorMap map_x1, const WgradTcParams p) {
  using S = WgradSmem<CWA, NA, CWB, TG, STAGES>;
  constexpr int SWA = CWA * 2, SWB = CWB * 2;
  constexpr uint32_t TMEM_COLS = (TG * CWB <= 32) ? 32 : (TG * CWB <= 64) ? 64 : (TG * CWB <= 128) ? 128 : 256;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SLACK);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.y -> (m tile, n tile, tap group)
  int by = blockIdx.y;
  const int tg = by % p.tap_groups; by /= p.tap_groups;
  const int n_tiles = p.n_tiles0 + p.n_tiles1;
  const int nt = by % n_tiles;
  const int mt = by / n_tiles;
  const int m0 = mt * 128;
  const bool src1 = nt >= p.n_tiles0;
  const int cb0 = (src1 ? nt - p.n_tiles0 : nt) * CWB;             // channel offset inside the source tensor
  const int ci_global = (src1 ? p.C0 : 0) + cb0;                   // channel offset in the concatenated input
  const int pad = p.ks >> 1;
  const int my_chunks = (p.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_dy);
    prefetch_tmap(src1 ? &map_x1 : &map_x0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mx = src1 ? &map_x1 : &map_x0;
      for (int it = 0; it < my_chunks; ++it) {
        const int chunk = blockIdx.x + it * gridDim.x;
        const int n = chunk / (p.tiles_x * p.tiles_y);
        const int tr = chunk - n * p.tiles_x * p.tiles_y;
        const int y0 = (tr / p.tiles_x) * TILE_H, x0 = (tr % p.tiles_x) * TILE_W;
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + NA * S::A_BOX;
        mbar_expect_tx(&full_bar[s], NA * S::A_BOX + TG * S::B_BOX);
#pragma unroll
        for (int i = 0; i < NA; ++i) tma_load_4d(&map_dy, &full_bar[s], a_dst + i * S::A_BOX, m0 + i * CWA, x0, y0, n);
#pragma unroll
        for (int tl = 0; tl < TG; ++tl) {
          const int t = tg * TG + tl;
          const int dy = (t / p.ks - pad) * p.dil, dx = (t % p.ks - pad) * p.dil;
          tma_load_4d(mx, &full_bar[s], b_dst + tl * S::B_BOX, cb0, x0 + dx, y0 + dy, n);
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_for(make_idesc_bf16(CWB, /*a_mn_major=*/1, /*b_mn_major=*/1, 128), p.dt);
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it % STAGES;
      const uint32_t ph = (it / STAGES) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + NA * S::A_BOX;
        const uint64_t adesc = make_mnmajor_desc<SWA>(a_addr, S::A_BOX);
#pragma unroll
        for (int tl = 0; tl < TG; ++tl) {
          const uint64_t bdesc = make_mnmajor_desc<SWB>(b_addr + tl * S::B_BOX, S::B_BOX);
#pragma unroll
          for (int k = 0; k < TILE_M / 16; ++k) {
            // 16 pixels (rows) per MMA: advance both operands by 16 rows
            umma_f16(tmem_base + tl * CWB, adesc + (uint64_t)((k * 16 * SWA) >> 4), bdesc + (uint64_t)((k * 16 * SWB) >> 4), idesc,
                     (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (it == my_chunks - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else if (my_chunks > 0) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int CinTot = p.C0 + p.C1;
    const float osc = (p.scale_a ? *p.scale_a : 1.f) * (p.scale_b ? *p.scale_b : 1.f);
    if (p.partials != nullptr) {          // deterministic split-K (see wgrad1_finalize_kernel)
      float* slot = p.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(TG * CWB * 128) + (q * 32 + lane);
#pragma unroll 1
      for (int c = 0; c < TG * CWB; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) slot[(size_t)(c + j) * 128] = v[j] * osc;
      }
    } else
#pragma unroll 1
    for (int tl = 0; tl < TG; ++tl) {
      const int t = tg * TG + tl;
#pragma unroll 1
      for (int c = 0; c < CWB; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(tl * CWB + c), v);
        if (row < p.CoutReal && q * 32 + lane < 128) {
          float* dst = p.dw + ((size_t)row * CinTot + ci_global + c) * p.taps + t;
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + (size_t)j * p.taps, v[j] * osc);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ================================================================================================
// wgrad v2 (3x3 only): one halo load of X per 8x16 pixel chunk; the nine tap operands are row-shifted MN-major
// descriptor views of that halo (same absolute-address swizzle argument as conv_tc2).  CTA = (Cout tile of 128) x
// (Cin block of CWB <= 32) x all nine taps -> nine accumulators of CWB columns in TMEM (<= 288 columns).
// L2->SM traffic per chunk: dY tile + 1.4x X block instead of dY + 9 X boxes.
// ================================================================================================
template <int CWA, int NA, int CWB, int STAGES>
struct Wgrad2Smem {
  static constexpr int A_BOX = TILE_M * CWA * 2;
  static constexpr int ROWB = CWB * 2;
  static constexpr int B_HALO = ((10 * 18 * ROWB + 1023) / 1024) * 1024;
  static constexpr int STAGE_BYTES = ((NA * A_BOX + B_HALO + 1023) / 1024) * 1024;
  static constexpr int SLACK = (128 / CWA) * A_BOX;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SLACK + 1024 + 256;
};

template <int SW>
__device__ __forceinline__ uint64_t make_mnmajor_desc_sbo(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  constexpr uint64_t layout = (SW == 128) ? 2 : (SW == 64) ? 4 : 6;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3ffff) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= layout << 61;
  return d;
}

template <int CWA, int NA, int CWB, int STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_tc2_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                   const __grid_constant__ CUtensorMap map_x0,
                                                                   const __grid_constant__ CUtensorMap map_x1, const WgradTcParams p) {
  using S = Wgrad2Smem<CWA, NA, CWB, STAGES>;
  constexpr int SWA = CWA * 2, SWB = CWB * 2;
  constexpr uint32_t TMEM_COLS = (9 * CWB <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SLACK);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_tiles0 + p.n_tiles1;
  const int nt = blockIdx.y % n_tiles;
  const int mt = blockIdx.y / n_tiles;
  const int m0 = mt * 128;
  const bool src1 = nt >= p.n_tiles0;
  const int cb0 = (src1 ? nt - p.n_tiles0 : nt) * CWB;
  const int ci_global = (src1 ? p.C0 : 0) + cb0;
  const int my_chunks = (p.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_dy);
    prefetch_tmap(src1 ? &map_x1 : &map_x0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mx = src1 ? &map_x1 : &map_x0;
      for (int it = 0; it < my_chunks; ++it) {
        const int chunk = blockIdx.x + it * gridDim.x;
        const int n = chunk / (p.tiles_x * p.tiles_y);
        const int tr = chunk - n * p.tiles_x * p.tiles_y;
        const int y0 = (tr / p.tiles_x) * 16, x0 = (tr % p.tiles_x) * 8;
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* a_dst = smem + s * S::STAGE_BYTES;
        uint8_t* b_dst = a_dst + NA * S::A_BOX;
        mbar_expect_tx(&full_bar[s], NA * S::A_BOX + 10 * 18 * S::ROWB);
#pragma unroll
        for (int i = 0; i < NA; ++i) tma_load_4d(&map_dy, &full_bar[s], a_dst + i * S::A_BOX, m0 + i * CWA, x0, y0, n);
        tma_load_4d(mx, &full_bar[s], b_dst, cb0, x0 - 1, y0 - 1, n);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_for(make_idesc_bf16(CWB, 1, 1, 128), p.dt);
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + NA * S::A_BOX;
        const uint64_t adesc = make_mnmajor_desc_sbo<SWA>(a_addr, S::A_BOX, 8 * SWA);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
          const int dy = t / 3, dx = t % 3;
          const uint64_t bdesc = make_mnmajor_desc_sbo<SWB>(b_addr + (uint32_t)((dy * 10 + dx) * S::ROWB), S::B_HALO, 10 * S::ROWB);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            // k-step = 16 pixels = image rows 2k, 2k+1 of the chunk: A advances 16 rows, B advances 2 halo rows
            umma_f16(tmem_base + t * CWB, adesc + (uint64_t)((k * 16 * SWA) >> 4), bdesc + (uint64_t)((k * 2 * 10 * S::ROWB) >> 4), idesc,
                     (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (it == my_chunks - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else if (my_chunks > 0) {
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int CinTot = p.C0 + p.C1;
#pragma unroll 1
    for (int t = 0; t < 9; ++t) {
#pragma unroll 1
      for (int c = 0; c < CWB; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * CWB + c), v);
        if (row < p.CoutReal) {
          float* dst = p.dw + ((size_t)row * CinTot + ci_global + c) * 9 + t;
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dst + (size_t)j * 9, v[j]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int CWA, int NA, int CWB>
int launch_wgrad2(const CUtensorMap& mdy, const CUtensorMap& mx0, const CUtensorMap& mx1, const WgradTcParams& p, int m_tiles,
                  cudaStream_t stream) {
  constexpr int per_stage = NA * TILE_M * CWA * 2 + 12 * 1024;
  constexpr int STAGES = per_stage >= 40 * 1024 ? 3 : 4;
  using S = Wgrad2Smem<CWA, NA, CWB, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc2_kernel<CWA, NA, CWB, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { wsl_set_error("wgrad_tc2: cudaFuncSetAttribute(%d bytes): %s", S::TOTAL, cudaGetErrorString(e)); return -5; }
    attr = true;
  }
  const int gy = m_tiles * (p.n_tiles0 + p.n_tiles1);
  int splits = (148 + gy - 1) / gy;
  if (splits > p.nchunks) splits = p.nchunks;
  if (splits < 1) splits = 1;
  dim3 grid(splits, gy);
  wgrad_tc2_kernel<CWA, NA, CWB, STAGES><<<grid, NUM_THREADS, S::TOTAL, stream>>>(mdy, mx0, mx1, p);
  return wsl_check_launch("wgrad_tc2");
}

// ================================================================================================
// wgrad v3 (3x3): taps in the M dimension.
//
//   A := X halo (MN-major).  An M = 128 operand is 128/CWB channel groups spaced LBO apart; with LBO = ONE PIXEL
//        (CWB*2 bytes) group g is the same halo viewed g pixels further right, i.e. filter column dx = g.  One MMA
//        therefore accumulates dx = 0,1,2 of a filter row at once (groups >= 3 are don't-care rows of D).
//   B := dY tile (MN-major), N = the Cout tile (<= 128) -> wide N keeps the tensor pipe ahead of the A smem reads.
//   D_dy[(dx, ci)][co], dy = 0..2: three accumulators of N columns in TMEM.
// MMAs per 128-pixel chunk and 32-channel Cin block: 3 dy x 8 k-steps = 24 (v2 issued 72 narrow ones).
// ================================================================================================
// DYN (NBOX == 1 only): the filter ROWS go into the MMA's N dimension as well.  dY is loaded with one halo row above and below
// (18 x 8 pixels); N-group g of the B operand is the dY tile viewed g - 1 rows further down (LBO = one tile row), while A is
// the un-shifted X row block (16 x 10 pixels, LBO = one pixel -> filter column in M).  X[p + (1 - g) rows + dx] * dY[p] lands
// in column group g, i.e. dy = 2 - g; the products that fall outside the chunk's rows are exactly those against the zero
// padding (TMA zero fill).  One N = 3*NTILE MMA replaces three N = NTILE ones: 8 instead of 24 per chunk; with the measured
// issue cost 10 + 32 (A read) + N/2 cycles that is 528 instead of 1224 cycles per chunk for the 16-channel layers.
template <int CWB, int CWN, int NBOX, int STAGES, bool DYN>
struct Wgrad3Smem {
  static constexpr int ROWB = CWB * 2;
  static constexpr int N_ROWS = DYN ? 18 : 16, X_ROWS = DYN ? 16 : 18;
  static constexpr int N_BOX = N_ROWS * 8 * CWN * 2;
  static constexpr int A_OFF = ((NBOX * N_BOX + 1023) / 1024) * 1024;
  static constexpr int A_BYTES = 10 * X_ROWS * ROWB;
  static constexpr int STAGE_BYTES = ((A_OFF + A_BYTES + 1023) / 1024) * 1024;
  static constexpr int SLACK = 2048;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + SLACK + 1024 + 256;
};

struct Wgrad3Params {
  int N, H, W;
  int C0, C1;
  int CoutReal;
  int tiles_x, tiles_y, nchunks;
  int k_tiles0, k_tiles1;     // Cin blocks (of CWB channels) in source 0 / 1
  int dt;                     // 16-bit operand type: 0 bf16, 2 fp16
  float* dw;
  float* partials;            // optional [gridDim.y][gridDim.x][3*NTILE][128]: per-CTA partial tiles for the fixed-order finalize
};

template <int CWB, int CWN, int NBOX, int STAGES, bool DYN>
__global__ void __launch_bounds__(NUM_THREADS, (DYN && CWN <= 32) ? 2 : 1) wgrad_tc3_kernel(const __grid_constant__ CUtensorMap map_dy,
                                                                   const __grid_constant__ CUtensorMap map_x0,
                                                                   const __grid_constant__ CUtensorMap map_x1, const Wgrad3Params p) {
  using S = Wgrad3Smem<CWB, CWN, NBOX, STAGES, DYN>;
  static_assert(!DYN || NBOX == 1, "row merging needs a single dY box");
  constexpr int NTILE = CWN * NBOX;
  constexpr int SWX = CWB * 2, SWN = CWN * 2;
  constexpr uint32_t TMEM_COLS = (3 * NTILE <= 64) ? 64 : (3 * NTILE <= 128) ? 128 : (3 * NTILE <= 256) ? 256 : 512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES + S::SLACK);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_tiles = p.k_tiles0 + p.k_tiles1;
  const int kt = blockIdx.y % k_tiles;          // Cin block
  const int nt = blockIdx.y / k_tiles;          // Cout tile
  const int n0 = nt * NTILE;
  const bool src1 = kt >= p.k_tiles0;
  const int cb0 = (src1 ? kt - p.k_tiles0 : kt) * CWB;
  const int ci_global = (src1 ? p.C0 : 0) + cb0;
  const int my_chunks = (p.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(accum_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_dy);
    prefetch_tmap(src1 ? &map_x1 : &map_x0);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      const CUtensorMap* mx = src1 ? &map_x1 : &map_x0;
      for (int it = 0; it < my_chunks; ++it) {
        const int chunk = blockIdx.x + it * gridDim.x;
        const int n = chunk / (p.tiles_x * p.tiles_y);
        const int tr = chunk - n * p.tiles_x * p.tiles_y;
        const int y0 = (tr / p.tiles_x) * 16, x0 = (tr % p.tiles_x) * 8;
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        uint8_t* b_dst = smem + s * S::STAGE_BYTES;          // dY boxes
        uint8_t* a_dst = b_dst + S::A_OFF;                   // X block
        mbar_expect_tx(&full_bar[s], NBOX * S::N_BOX + S::A_BYTES);
#pragma unroll
        for (int i = 0; i < NBOX; ++i)
          tma_load_4d(&map_dy, &full_bar[s], b_dst + i * S::N_BOX, n0 + i * CWN, x0, DYN ? y0 - 1 : y0, n);
        tma_load_4d(mx, &full_bar[s], a_dst, cb0, x0 - 1, DYN ? y0 : y0 - 1, n);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_for(make_idesc_bf16(DYN ? 3 * NTILE : NTILE, 1, 1, 128), p.dt);
    // A: channel-group stride (LBO) = one pixel, so group g == filter column dx = g; 8-pixel-row groups 10 halo pixels apart
    constexpr uint32_t a_hi = desc_hi<SWX>(10 * S::ROWB), b_hi = desc_hi<SWN>(8 * SWN);
    constexpr uint32_t a_lbo = ((uint32_t)(S::ROWB >> 4) & 0x3fffu) << 16;
    constexpr uint32_t b_lbo = ((uint32_t)((DYN ? 8 * SWN : S::N_BOX) >> 4) & 0x3fffu) << 16;
    const uint32_t elected = elect_one();
    const uint32_t tmem_u = uniform(tmem_base);
    const uint32_t s_lo0 = uniform(smem_u32(smem)) >> 4;
    for (int it = 0; it < my_chunks; ++it) {
      const int s = it % STAGES;
      mbar_wait(&full_bar[s], (it / STAGES) & 1);
      tc_fence_after();
      if (elected) {
        const uint32_t b_lo = (s_lo0 + (uint32_t)(s * (S::STAGE_BYTES >> 4))) | b_lbo;
        const uint32_t a_lo = (s_lo0 + (uint32_t)(s * (S::STAGE_BYTES >> 4) + (S::A_OFF >> 4))) | a_lbo;
        if constexpr (DYN) {
#pragma unroll
          for (int k = 0; k < 8; ++k)      // k-step = 16 pixels = chunk rows 2k, 2k+1
            umma_f16_w(tmem_u, a_lo + (uint32_t)((k * 2 * 10 * S::ROWB) >> 4), a_hi, b_lo + (uint32_t)((k * 16 * SWN) >> 4), b_hi, idesc,
                       (it > 0 || k > 0) ? 1u : 0u);
        } else {
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_f16_w(tmem_u + (uint32_t)(dy * NTILE), a_lo + (uint32_t)(((dy * 10 * S::ROWB) >> 4) + ((k * 2 * 10 * S::ROWB) >> 4)), a_hi,
                         b_lo + (uint32_t)((k * 16 * SWN) >> 4), b_hi, idesc, (it > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[s]);
        if (it == my_chunks - 1) umma_commit(accum_bar);
      }
      __syncwarp();
    }
  } else if (my_chunks > 0) {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int g = m / CWB, ci = m % CWB;        // filter column dx, input channel inside the block
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int CinTot = p.C0 + p.C1;
    if (p.partials != nullptr) {
      // deterministic split-K: this CTA's partial tile goes to its own slot (coalesced: lane = TMEM lane is the fastest index);
      // wgrad3_finalize_kernel sums the slots of an output tile in split order and adds the result into dw
      float* slot = p.partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (size_t)(3 * NTILE * 128) + m;
#pragma unroll 1
      for (int c = 0; c < 3 * NTILE; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) slot[(size_t)(c + j) * 128] = v[j];
      }
    } else
#pragma unroll 1
    for (int grp = 0; grp < 3; ++grp) {
      const int dy = DYN ? 2 - grp : grp;       // column group -> filter row
#pragma unroll 1
      for (int c = 0; c < NTILE; c += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(grp * NTILE + c), v);
        if (g < 3) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int co = n0 + c + j;
            if (co < p.CoutReal) atomicAdd(p.dw + ((size_t)co * CinTot + ci_global + ci) * 9 + dy * 3 + g, v[j]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// Fixed-order reduction of the per-CTA partial tiles of one wgrad_tc3 launch: element e = (grp * NTILE + col) * 128 + m of output tile
// blockIdx.y is summed over the `splits` slots in index order (independent of which CTA finished when) and ADDED to the torch-layout
// gradient (accumulation semantics are kept: a second backward through shared weights adds on top).
__global__ void __launch_bounds__(256) wgrad3_finalize_kernel(const float* __restrict__ partials, int splits, int ntile, int cwb, int dyn,
                                                              int k_tiles0, int k_tiles, int C0, int CinTot, int CoutReal,
                                                              float* __restrict__ dw) {
  const int E = 3 * ntile * 128;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= E) return;
  const int m = e & 127, col = (e >> 7) % ntile, grp = (e >> 7) / ntile;
  const int g = m / cwb, ci = m % cwb;
  const int kt = blockIdx.y % k_tiles, nt = blockIdx.y / k_tiles;
  const int co = nt * ntile + col;
  if (g >= 3 || co >= CoutReal) return;
  const bool src1 = kt >= k_tiles0;
  const int ci_global = (src1 ? C0 : 0) + (src1 ? kt - k_tiles0 : kt) * cwb;
  const float* src = partials + (size_t)blockIdx.y * splits * E + e;
  float acc = 0.f;
  int s = 0;
  for (; s + 8 <= splits; s += 8) {
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (size_t)(s + u) * E);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; s < splits; ++s) acc += __ldcg(src + (size_t)s * E);
  const int dy = dyn ? 2 - grp : grp;
  dw[((size_t)co * CinTot + ci_global + ci) * 9 + dy * 3 + g] += acc;
}

template <int CWB, int CWN, int NBOX>
int launch_wgrad3(const CUtensorMap& mdy, const CUtensorMap& mx0, const CUtensorMap& mx1, Wgrad3Params p, int n_tiles,
                  float* partial_ws, long long partial_floats, cudaStream_t stream) {
  constexpr bool DYN = (NBOX == 1);
  constexpr int per_stage = NBOX * TILE_M * CWN * 2 + 12 * 1024;
  constexpr int STAGES = per_stage >= 40 * 1024 ? 3 : 4;
  using S = Wgrad3Smem<CWB, CWN, NBOX, STAGES, DYN>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc3_kernel<CWB, CWN, NBOX, STAGES, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { wsl_set_error("wgrad_tc3: cudaFuncSetAttribute(%d bytes): %s", S::TOTAL, cudaGetErrorString(e)); return -5; }
    attr = true;
  }
  const int gy = n_tiles * (p.k_tiles0 + p.k_tiles1);
  // narrow layers are bounded by the per-CTA rate of 32/64-byte TMA rows: two CTAs per SM (TMEM and smem allow it)
  constexpr int per_sm = (DYN && CWN <= 32 && S::TOTAL <= 100 * 1024) ? 2 : 1;
  int splits = (148 * per_sm + gy - 1) / gy;
  if (splits > p.nchunks) splits = p.nchunks;
  if (splits < 1) splits = 1;
  dim3 grid(splits, gy);
  constexpr int NTILE = CWN * NBOX;
  const long long need = (long long)splits * gy * 3 * NTILE * 128;
  p.partials = (partial_ws != nullptr && need <= partial_floats) ? partial_ws : nullptr;     // too small a workspace: atomics
  wgrad_tc3_kernel<CWB, CWN, NBOX, STAGES, DYN><<<grid, NUM_THREADS, S::TOTAL, stream>>>(mdy, mx0, mx1, p);
  int rc = wsl_check_launch("wgrad_tc3");
  if (rc || p.partials == nullptr) return rc;
  wgrad3_finalize_kernel<<<dim3((3 * NTILE * 128 + 255) / 256, gy), 256, 0, stream>>>(p.partials, splits, NTILE, CWB, DYN ? 1 : 0, p.k_tiles0,
                                                                                     p.k_tiles0 + p.k_tiles1, p.C0, p.C0 + p.C1, p.CoutReal, p.dw);
  return wsl_check_launch("wgrad3_finalize");
}

// fixed-order reduction of wgrad_tc_kernel's per-CTA partial tiles: element e = (tl * CWB + c) * 128 + lane of output tile blockIdx.y
__global__ void __launch_bounds__(256) wgrad1_finalize_kernel(const float* __restrict__ partials, int splits, int tg_size, int cwb,
                                                              int tap_groups, int n_tiles0, int n_tiles, int C0, int CinTot,
                                                              int CoutReal, int taps, float* __restrict__ dw) {
  const int E = tg_size * cwb * 128;
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= E) return;
  const int lane = e & 127, c = (e >> 7) % cwb, tl = (e >> 7) / cwb;
  int by = blockIdx.y;
  const int tg = by % tap_groups; by /= tap_groups;
  const int nt = by % n_tiles, mt = by / n_tiles;
  const int row = mt * 128 + lane;
  if (row >= CoutReal) return;
  const bool src1 = nt >= n_tiles0;
  const int ci = (src1 ? C0 : 0) + (src1 ? nt - n_tiles0 : nt) * cwb + c;
  const float* src = partials + (size_t)blockIdx.y * splits * E + e;
  float acc = 0.f;
  int s = 0;
  for (; s + 8 <= splits; s += 8) {            // eight loads in flight, fixed summation order
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (size_t)(s + u) * E);
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += v[u];
  }
  for (; s < splits; ++s) acc += __ldcg(src + (size_t)s * E);
  dw[((size_t)row * CinTot + ci) * taps + tg * tg_size + tl] += acc;
}

static thread_local float* g_wgrad1_ws = nullptr;        // set by the entry point for the launch helpers below
static thread_local long long g_wgrad1_ws_floats = 0;

template <int CWA, int NA, int CWB, int TG>
int launch_wgrad(const CUtensorMap& mdy, const CUtensorMap& mx0, const CUtensorMap& mx1, WgradTcParams p, int m_tiles,
                 cudaStream_t stream) {
  constexpr int per_stage = NA * TILE_M * CWA * 2 + TG * TILE_M * CWB * 2;
  constexpr int STAGES = per_stage >= 80 * 1024 ? 2 : per_stage >= 48 * 1024 ? 3 : 4;
  using S = WgradSmem<CWA, NA, CWB, TG, STAGES>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<CWA, NA, CWB, TG, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) { wsl_set_error("wgrad_tc: cudaFuncSetAttribute(%d bytes): %s", S::TOTAL, cudaGetErrorString(e)); return -5; }
    attr = true;
  }
  const int gy = m_tiles * (p.n_tiles0 + p.n_tiles1) * p.tap_groups;
  int splits = (148 * 2 + gy - 1) / gy;
  if (splits > p.nchunks) splits = p.nchunks;
  if (splits < 1) splits = 1;
  dim3 grid(splits, gy);
  const long long need = (long long)splits * gy * TG * CWB * 128;
  p.partials = (g_wgrad1_ws != nullptr && need <= g_wgrad1_ws_floats) ? g_wgrad1_ws : nullptr;
  wgrad_tc_kernel<CWA, NA, CWB, TG, STAGES><<<grid, NUM_THREADS, S::TOTAL, stream>>>(mdy, mx0, mx1, p);
  int rc = wsl_check_launch("wgrad_tc");
  if (rc || p.partials == nullptr) return rc;
  wgrad1_finalize_kernel<<<dim3((TG * CWB * 128 + 255) / 256, gy), 256, 0, stream>>>(p.partials, splits, TG, CWB, p.tap_groups, p.n_tiles0,
                                                                                   p.n_tiles0 + p.n_tiles1, p.C0, p.C0 + p.C1, p.CoutReal,
                                                                                   p.taps, p.dw);
  return wsl_check_launch("wgrad1_finalize");
}

template <int CWA, int NA, int CWB>
int wgrad_dispatch_tg(const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const WgradTcParams& p, int m_tiles, cudaStream_t st) {
  if (p.ks == 3) return launch_wgrad<CWA, NA, CWB, 3>(a, b0, b1, p, m_tiles, st);
  return launch_wgrad<CWA, NA, CWB, 1>(a, b0, b1, p, m_tiles, st);
}
template <int CWA, int NA>
int wgrad_dispatch_b(int cwb, const CUtensorMap& a, const CUtensorMap& b0, const CUtensorMap& b1, const WgradTcParams& p, int m_tiles, cudaStream_t st) {
  switch (cwb) {
    case 64: return wgrad_dispatch_tg<CWA, NA, 64>(a, b0, b1, p, m_tiles, st);
    case 32: return wgrad_dispatch_tg<CWA, NA, 32>(a, b0, b1, p, m_tiles, st);
    default: return wgrad_dispatch_tg<CWA, NA, 16>(a, b0, b1, p, m_tiles, st);
  }
}

// per-channel sum of a channels-last bf16 tensor, added into out[C] (bias gradients of conv1x1 / out_conv)
template <typename T, bool HALF = false>
__global__ void __launch_bounds__(256) channel_sum_kernel(const T* __restrict__ x, long long P, int C, int Creal,
                                                          float* __restrict__ out, float* __restrict__ ws) {
  extern __shared__ float s_red[];
  const int cg = C >> 3, rows = 256 / cg;
  const int g = threadIdx.x % cg, r = threadIdx.x / cg;
  float sum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sum[j] = 0.f;
  for (long long q = (long long)blockIdx.x * rows + r; r < rows && q < P; q += (long long)gridDim.x * rows) {   // threads beyond rows*cg idle
    float v[8];
    if constexpr (sizeof(T) == 2) {
      unpack8_dt(*reinterpret_cast<const uint4*>(x + q * C + g * 8), v, HALF ? 2 : 0);
    } else {
      const float4 a = reinterpret_cast<const float4*>(x + q * C + g * 8)[0], b = reinterpret_cast<const float4*>(x + q * C + g * 8)[1];
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] += v[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) s_red[threadIdx.x * 8 + j] = sum[j];
  __syncthreads();
  if (ws == nullptr) {
    for (int c = threadIdx.x; c < Creal; c += 256) {
      float a = 0.f;
      for (int rr = 0; rr < rows; ++rr) a += s_red[(rr * cg + (c >> 3)) * 8 + (c & 7)];
      atomicAdd(out + c, a);
    }
    return;
  }
  // deterministic: per-block rows, the last block to finish (ticket in ws[0]) adds them up in block order
  for (int c = threadIdx.x; c < Creal; c += 256) {
    float a = 0.f;
    for (int rr = 0; rr < rows; ++rr) a += s_red[(rr * cg + (c >> 3)) * 8 + (c & 7)];
    ws[64 + (size_t)blockIdx.x * Creal + c] = a;
  }
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(reinterpret_cast<unsigned*>(ws), 1u) == gridDim.x - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // all 256 threads: `parts` threads per channel walk interleaved block subsets (8 loads in flight), then a fixed-order combine;
  // more than 256 channels (PNet2D's 320-channel concat head) go in slabs of 256
  __shared__ float s_fin[256];
  for (int cb = 0; cb < Creal; cb += 256) {
    const int cn = Creal - cb < 256 ? Creal - cb : 256;
    const int parts = 256 / cn;
    const int c = cb + threadIdx.x % cn, part = threadIdx.x / cn;
    float a = 0.f;
    if (part < parts) {
      int b = part;
      for (; b + 7 * parts < (int)gridDim.x; b += 8 * parts) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcg(&ws[64 + (size_t)(b + u * parts) * Creal + c]);
#pragma unroll
        for (int u = 0; u < 8; ++u) a += v[u];
      }
      for (; b < (int)gridDim.x; b += parts) a += __ldcg(&ws[64 + (size_t)b * Creal + c]);
    }
    s_fin[threadIdx.x] = a;
    __syncthreads();
    if (threadIdx.x < cn) {
      float t = 0.f;
      for (int q = 0; q < parts; ++q) t += s_fin[q * cn + threadIdx.x];
      out[cb + threadIdx.x] += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *reinterpret_cast<unsigned*>(ws) = 0u;
}


static int g_conv2_last_rows = 0;   // gridDim.x of the most recent conv_tc2 launch (rows of stat_partials)

template <int KS, int KBLK, int NT, int MT>
int launch_conv2(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvV2Params& p, cudaStream_t stream) {
  using Cfg = ConvV2Cfg<KS, KBLK, NT, MT>;
  constexpr int STAGES = (Cfg::A_STAGE <= 8 * 1024) ? 4 : (Cfg::A_STAGE <= 13 * 1024) ? 3 : 2;
  const int T = KS * KS;
  const int nkb = (p.C0 + p.C1) / KBLK;
  const int w_bytes = ((T * nkb * Cfg::W_SUB + 1023) / 1024) * 1024;
  const int smem = w_bytes + STAGES * Cfg::A_STAGE + 1024 + 256;
  if (smem > 227 * 1024) { wsl_set_error("conv_tc2: %d bytes of shared memory needed", smem); return -6; }
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<KS, KBLK, NT, MT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { wsl_set_error("conv_tc2: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return -5; }
    attr_bytes = smem;
  }
  const int n_tiles = p.CoutP / NT;
  int occ = (220 * 1024) / smem;
  const int occ_tmem = 512 / (int)Cfg::TMEM_COLS;
  if (occ > occ_tmem) occ = occ_tmem;
  if (occ > 4) occ = 4;
  if (occ < 1) occ = 1;
  int gx = (148 * occ) / n_tiles;
  if (gx < 1) gx = 1;
  if (gx > p.ntiles) gx = p.ntiles;
  dim3 grid(gx, n_tiles);
  g_conv2_last_rows = gx;
  conv_tc2_kernel<KS, KBLK, NT, MT, STAGES><<<grid, Conv2Epi<NT, MT>::THREADS, smem, stream>>>(a0, a1, b, p, w_bytes);
  return wsl_check_launch("conv_tc2");
}

template <int KS, int KBLK, int MT>
int conv2_dispatch_nt(int nt, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvV2Params& p, cudaStream_t st) {
  switch (nt) {
    case 16:  return launch_conv2<KS, KBLK, 16, MT>(a0, a1, b, p, st);
    case 32:  return launch_conv2<KS, KBLK, 32, MT>(a0, a1, b, p, st);
    case 64:  return launch_conv2<KS, KBLK, 64, MT>(a0, a1, b, p, st);
    case 128: if constexpr (MT <= 2) return launch_conv2<KS, KBLK, 128, MT>(a0, a1, b, p, st); else break;
  }
  wsl_set_error("conv_tc2: unsupported N tile %d", nt);
  return -1;
}


// ================================================================================================
// conv_row: 3x3 convolution for narrow outputs (CoutP = 16 or 32) with the filter ROWS in the MMA's N dimension.
//
// A tcgen05.mma (SS, M = 128, K = 16) costs ~43 + N/2 cycles on the B200 (profiles/micro/mma_rate.cu): the 4 KB A-operand
// read dominates when N = Cout is 16 or 32, and conv_tc2's nine N = 16 MMAs per 128 pixels (459 cycles) bound the
// full-resolution layers, not HBM.  Here the work unit is one INPUT row of 128 pixels:
//     E_i[c][(dy, co)] = sum_{dx, ci} X[row i][c + dx - 1][ci] * W[dy][dx][co][ci]
// i.e. 3 MMAs per 16 input channels (dx = pixel-shifted views of ONE row buffer), each with N = 3*Cout, accumulated into a
// ring slot of 3*Cout TMEM columns; an output row is then
//     out[r][c] = E_r[c][dy=0] + E_{r+1}[c][dy=1] + E_{r+2}[c][dy=2]
// -- three column groups of three ring slots in the SAME TMEM lane, summed by the epilogue thread that owns pixel c: no
// cross-lane traffic.  201 cycles per 128 pixels for 16 -> 16 channels instead of 459.
//   warp 0: TMA producer (one [16 ch x 130 px] box per input row and 16-channel block; OOB zero fill = padding)
//   warp 1: MMA issuer;  warps 2..5: epilogue (lane = pixel; 1 KB contiguous bf16 per warp and output row).
// Work item = (image, 128-pixel column strip, segment of RS output rows); persistent CTAs walk the items.
// ================================================================================================
constexpr int ROW_PX = 128;
// bytes of one input-row buffer: 130 halo pixels x KBW channels, padded to a multiple of the 8-row swizzle atom
__host__ __device__ constexpr int row_abytes(int kbw) { return ((130 * kbw * 2 + 255) / 256) * 256; }

template <int NT, int RING, int EW, int KBW>
__global__ void __launch_bounds__(64 + 32 * EW, (RING * 3 * NT <= 256) ? 2 : 1) conv_row_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                                                       const __grid_constant__ CUtensorMap map_b, const ConvV2Params p, const int KB,
                                                       const int RS, const int w_bytes, const int STAGES) {
  constexpr int NG = 3 * NT;                              // accumulator columns of one input row (three filter rows)
  constexpr int SWB = KBW * 2, ROW_ABYTES = row_abytes(KBW);   // channel block = swizzle span: 32 B (16 ch) or 64 B (32 ch)
  // (a 64-byte block halves the number of TMA rows per box: the producer is bounded by ~5 cycles per TMA row and CTA)
  constexpr uint32_t TMEM_COLS = (RING * NG <= 256) ? 256 : 512;
  static_assert(RING * NG <= 512 && RING >= 4, "TMEM ring");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* w_smem = smem;                                 // [KB][dx][3*NT rows x 32 B]
  uint8_t* a_smem = smem + w_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(a_smem + (size_t)STAGES * KB * ROW_ABYTES);
  uint64_t* a_full = bars;
  uint64_t* a_empty = bars + STAGES;
  uint64_t* e_full = bars + 2 * STAGES;
  uint64_t* e_empty = e_full + RING;
  uint64_t* w_bar = e_empty + RING;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = p.C0 / KBW;
  const int strips = p.W / ROW_PX, segs = p.H / RS;
  const int items = p.N * strips * segs;
  const int my_items = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int rows_in = RS + 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < RING; ++s) { mbar_init(&e_full[s], 1); mbar_init(&e_empty[s], 4); }
    mbar_init(w_bar, 1);
    fence_barrier_init();
    prefetch_tmap(&map_a0);
    prefetch_tmap(&map_b);
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(w_bar, (uint32_t)(KB * 9 * NT * SWB));
      for (int kb = 0; kb < KB; ++kb)
        for (int dx = 0; dx < 3; ++dx)
          for (int dy = 0; dy < 3; ++dy)
            tma_load_3d(&map_b, w_bar, w_smem + (((kb * 3 + dx) * 3 + dy) * NT) * SWB, kb * KBW, 0, dy * 3 + dx);
    }
    __syncwarp();
    // row loop: the warp stays converged, one elected lane issues with warp-uniform operands (a divergent lane-0 branch
    // costs a few hundred cycles per TMA instruction, and this kernel issues one or two per 128 pixels)
    const uint32_t elected = elect_one();
    const uint32_t a_u = uniform(smem_u32(a_smem)), full_u = uniform(smem_u32(a_full));
    int s = 0;
    uint32_t sph = 1;                    // parity to wait for on a_empty[s] (fresh barrier: passes)
    for (int it = 0; it < my_items; ++it) {
      const int item = blockIdx.x + it * gridDim.x;
      const int seg = item % segs, t1 = item / segs, strip = t1 % strips, n = t1 / strips;
      const int y0 = seg * RS, x0 = strip * ROW_PX;
      for (int i = 0; i < rows_in; ++i) {
        mbar_wait(&a_empty[s], sph);
        if (elected) {
          const uint32_t bar = full_u + (uint32_t)(s * 8);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)(KB * 130 * SWB)) : "memory");
          for (int kb = 0; kb < KB; ++kb) {
            const uint32_t dst = a_u + (uint32_t)((s * KB + kb) * ROW_ABYTES);
            const CUtensorMap* mp = (kb < kb0) ? &map_a0 : &map_a1;
            const int c0 = (kb < kb0 ? kb : kb - kb0) * KBW;
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(dst), "l"((uint64_t)mp), "r"(bar), "r"(c0), "r"(x0 - 1), "r"(y0 - 1 + i), "r"(n)
                : "memory");
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; sph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_for(make_idesc_bf16(NG), p.dt);
    constexpr uint32_t d_hi = desc_hi<SWB>(8 * SWB);
    const uint32_t elected = elect_one();
    const uint32_t tmem_u = uniform(tmem_base);
    const uint32_t a_lo0 = (uniform(smem_u32(a_smem)) >> 4) | 0x10000u;
    const uint32_t w_lo0 = (uniform(smem_u32(w_smem)) >> 4) | 0x10000u;
    mbar_wait(w_bar, 0);
    const int total_rows = my_items * rows_in;
    int s = 0, slot = 0;
    uint32_t sph = 0, rph = 1;           // parities: a_full[s] (first completion), e_empty[slot] (fresh barrier passes)
    for (int q = 0; q < total_rows; ++q) {
      mbar_wait(&e_empty[slot], rph);
      mbar_wait(&a_full[s], sph);
      tc_fence_after();
      if (elected) {
        const uint32_t d = tmem_u + (uint32_t)(slot * NG);
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t a_lo = a_lo0 + (uint32_t)((s * KB + kb) * (ROW_ABYTES >> 4));
          const uint32_t w_lo = w_lo0 + (uint32_t)(kb * 3 * ((NG * SWB) >> 4));
#pragma unroll
          for (int ks = 0; ks < KBW / 16; ++ks) {
#pragma unroll
            for (int dx = 0; dx < 3; ++dx)
              umma_f16_w(d, a_lo + (uint32_t)(dx * (SWB >> 4) + ks * 2), d_hi, w_lo + (uint32_t)(dx * ((NG * SWB) >> 4) + ks * 2), d_hi, idesc,
                         (kb > 0 || ks > 0 || dx > 0) ? 1u : 0u);
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&e_full[slot]);
      }
      __syncwarp();
      if (++s == STAGES) { s = 0; sph ^= 1u; }
      if (++slot == RING) { slot = 0; rph ^= 1u; }
    }
  } else {
    // EW epilogue warps: warp%4 selects the TMEM lane quarter (= 32-pixel group of the strip), (warp-2)/4 the set of output
    // rows (j % SETS) it owns.  The per-128-pixel MMA time is short here, so the epilogue must be cheap: bias from shared
    // memory, BatchNorm partial sums accumulated per lane (pixel) and transposed once at the end.
    constexpr int SETS = EW / 4;
    __shared__ float s_bias[NT];
    __shared__ float s_stat[EW][2][NT];
    const int ew = warp - 2, qt = warp & 3, set = ew >> 2;
    if (ew == 0 && lane < NT) s_bias[lane] = p.bias ? p.bias[lane] : 0.f;
    asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
    const uint32_t lane_base = tmem_base + ((uint32_t)(qt * 32) << 16);
    float ls[NT], lq[NT];
#pragma unroll
    for (int c = 0; c < NT; ++c) ls[c] = lq[c] = 0.f;
    int q0 = 0;
    for (int it = 0; it < my_items; ++it, q0 += rows_in) {
      const int item = blockIdx.x + it * gridDim.x;
      const int seg = item % segs, t1 = item / segs, strip = t1 % strips, n = t1 / strips;
      const int y0 = seg * RS, gx = strip * ROW_PX + qt * 32 + lane;
      for (int j = set; j < RS; j += SETS) {
        const int qa = q0 + j, qb = qa + 1, qc = qa + 2;
        mbar_wait(&e_full[qc % RING], (qc / RING) & 1);    // commits complete in order: rows qa, qb are done as well
        tc_fence_after();
        const int gy = y0 + j;
#pragma unroll
        for (int c = 0; c < NT; c += 16) {
          float v[16], u1[16], u2[16];
          tmem_ld16_nowait(lane_base + (uint32_t)((qa % RING) * NG + 0 * NT + c), v);
          tmem_ld16_nowait(lane_base + (uint32_t)((qb % RING) * NG + 1 * NT + c), u1);
          tmem_ld16_nowait(lane_base + (uint32_t)((qc % RING) * NG + 2 * NT + c), u2);
          tmem_ld_wait();
          if (c + 16 >= NT) {                               // all reads of this output row are done: release ring slots
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              mbar_arrive(&e_empty[qa % RING]);
              if (j == RS - 1) { mbar_arrive(&e_empty[qb % RING]); mbar_arrive(&e_empty[qc % RING]); }
            }
          }
#pragma unroll
          for (int jj = 0; jj < 16; jj += 4) {
            const float4 bb = *reinterpret_cast<const float4*>(&s_bias[c + jj]);
            v[jj + 0] = ((v[jj + 0] + u1[jj + 0]) + u2[jj + 0]) + bb.x;
            v[jj + 1] = ((v[jj + 1] + u1[jj + 1]) + u2[jj + 1]) + bb.y;
            v[jj + 2] = ((v[jj + 2] + u1[jj + 2]) + u2[jj + 2]) + bb.z;
            v[jj + 3] = ((v[jj + 3] + u1[jj + 3]) + u2[jj + 3]) + bb.w;
          }
          if (p.out_mode == 0) {
            float lo[8], hi[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) { lo[jj] = v[jj]; hi[jj] = v[8 + jj]; }
            const uint4 plo = pack8_dt(lo, p.dt), phi = pack8_dt(hi, p.dt);
            if (c < p.CoutStore) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (((long long)n * p.H + gy) * p.W + gx) * p.CoutStore + c;
              reinterpret_cast<uint4*>(o)[0] = plo;
              if (c + 8 < p.CoutStore) reinterpret_cast<uint4*>(o)[1] = phi;
            }
            if (p.stat_partials) {   // statistics of the values exactly as stored (rounded to the 16-bit type)
              unpack8_dt(plo, lo, p.dt);
              unpack8_dt(phi, hi, p.dt);
#pragma unroll
              for (int jj = 0; jj < 8; ++jj) {
                ls[c + jj] += lo[jj];           lq[c + jj] = fmaf(lo[jj], lo[jj], lq[c + jj]);
                ls[c + 8 + jj] += hi[jj];       lq[c + 8 + jj] = fmaf(hi[jj], hi[jj], lq[c + 8 + jj]);
              }
            }
          } else {
            float* o = reinterpret_cast<float*>(p.out);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj)
              if (c + jj < p.CoutStore) o[(((long long)n * p.CoutStore + c + jj) * p.H + gy) * p.W + gx] = v[jj];
          }
        }
      }
      // every set must have released its rows before the ring indices of the next item are reused: nothing to do, the
      // e_empty arrivals above are per row and each row is owned by exactly one set
    }
    if (p.stat_partials) {
#pragma unroll
      for (int c = 0; c < NT; c += 16) {
        float a[16], b[16];
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) { a[jj] = ls[c + jj]; b[jj] = lq[c + jj]; }
        const float sa = warp_transpose_sum16(a, lane), sb = warp_transpose_sum16(b, lane);
        if ((lane & 1) == 0) {
          s_stat[ew][0][c + ((lane >> 1) & 15)] = sa;
          s_stat[ew][1][c + ((lane >> 1) & 15)] = sb;
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * EW) : "memory");
      const int t = threadIdx.x - 64;
      for (int c = t; c < 2 * NT; c += 32 * EW) {
        const int which = c / NT, ch = c % NT;
        float a = 0.f;
#pragma unroll
        for (int w = 0; w < EW; ++w) a += s_stat[w][which][ch];
        p.stat_partials[((size_t)blockIdx.x * 2 + which) * p.CoutP + ch] = a;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NT, int RING, int EW, int KBW>
int launch_conv_row_cfg(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvV2Params& p, int KB, int RS,
                        int STAGES, cudaStream_t stream) {
  const int w_bytes = ((KB * 9 * NT * KBW * 2 + 1023) / 1024) * 1024;
  const int smem = w_bytes + STAGES * KB * row_abytes(KBW) + 1024 + 256;
  if (smem > 220 * 1024) { wsl_set_error("conv_row: %d bytes of shared memory needed", smem); return -6; }
  static int attr_bytes = 0;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(conv_row_kernel<NT, RING, EW, KBW>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { wsl_set_error("conv_row: cudaFuncSetAttribute(%d): %s", smem, cudaGetErrorString(e)); return -5; }
    attr_bytes = smem;
  }
  const int items = p.N * (p.W / ROW_PX) * (p.H / RS);
  int occ = (RING * 3 * NT <= 256) ? 2 : 1;
  if ((220 * 1024) / smem < occ) occ = 1;
  int gx = 148 * occ;
  if (gx > items) gx = items;
  g_conv2_last_rows = gx;
  conv_row_kernel<NT, RING, EW, KBW><<<gx, 64 + 32 * EW, smem, stream>>>(a0, a1, b, p, KB, RS, w_bytes, STAGES);
  return wsl_check_launch("conv_row");
}

template <int NT>
int launch_conv_row(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b, const ConvV2Params& p, int KB, int RS, int kbw,
                    cudaStream_t stream) {
  // NT = 16: either one CTA per SM owning the whole TMEM (ring of 10 row accumulators, 12 epilogue warps) or two CTAs with
  // half of it each (ring of 5, 8 epilogue warps); NT = 32: one CTA (ring of 5 x 96 columns), 8 epilogue warps.
  // KB counts channel blocks of kbw channels.  Row buffers in flight: ~80 KB per SM cover the HBM latency.
  static const int cfg = [] { const char* e = getenv("WSL4MIS_ROW_CFG"); return e ? atoi(e) : 1; }();
  static const int st_env = [] { const char* e = getenv("WSL4MIS_ROW_STAGES"); return e ? atoi(e) : 0; }();
  if constexpr (NT == 16) {
    if (cfg == 1 && KB <= 2) return launch_conv_row_cfg<16, 5, 8, 16>(a0, a1, b, p, KB, RS, st_env ? st_env : (KB == 1 ? 10 : 6), stream);
    return launch_conv_row_cfg<16, 10, 12, 16>(a0, a1, b, p, KB, RS, st_env ? st_env : (KB <= 2 ? 12 : 8), stream);
  } else {
    if (kbw == 32) return launch_conv_row_cfg<32, 5, 8, 32>(a0, a1, b, p, KB, RS, st_env ? st_env : (KB == 1 ? 12 : 8), stream);
    return launch_conv_row_cfg<32, 5, 8, 16>(a0, a1, b, p, KB, RS, st_env ? st_env : (KB == 1 ? 16 : KB == 2 ? 12 : 8), stream);
  }
}

}  // namespace

WSL_API int wsl_tc_available(void) { return get_encode() != nullptr ? 1 : 0; }

static int conv_tc_impl(const void* src0, int C0, const void* src1, int C1, const void* wpk_bf16, const float* bias, void* out,
                        int out_mode, int N, int H, int W, int CoutP, int CoutStore, int ksize, int dtype, int dilation,
                        cudaStream_t stream) {
  WSL_REQUIRE(ksize == 3 || ksize == 1, "wsl_conv_tc: ksize must be 1 or 3");
  WSL_REQUIRE(dilation >= 1 && dilation <= 64, "wsl_conv_tc: dilation must be in 1..64 (got %d)", dilation);
  WSL_REQUIRE(dtype == 0 || dtype == 2, "wsl_conv_tc: dtype must be 0 (bf16) or 2 (fp16)");
  WSL_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "wsl_conv_tc: source channels must be multiples of 16 (got %d,%d)", C0, C1);
  WSL_REQUIRE(CoutP % 16 == 0, "wsl_conv_tc: CoutP must be a multiple of 16");
  WSL_REQUIRE(W % TILE_W == 0 && H % TILE_H == 0, "wsl_conv_tc: H,W must be multiples of the 8x16 pixel tile (got %dx%d)", H, W);
  // k-block = largest of 64/32/16 dividing both sources' channel counts
  int kblk = 64;
  while (kblk > 16 && (C0 % kblk != 0 || (C1 > 0 && C1 % kblk != 0))) kblk >>= 1;
  int nt = CoutP >= 128 ? 128 : CoutP;
  while (nt > 16 && CoutP % nt != 0) nt >>= 1;          // e.g. CoutP = 320 (PNet2D's concat block): five tiles of 64
  WSL_REQUIRE(CoutP % nt == 0 && (nt == 16 || nt == 32 || nt == 64 || nt == 128), "wsl_conv_tc: unsupported CoutP %d", CoutP);
  const int CinP = C0 + C1, T = ksize * ksize;
  CUtensorMap a0, a1, b;
  {
    long long d[4] = {C0, W, H, N};
    int bx[4] = {kblk, TILE_W, TILE_H, 1};
    int rc = get_map(src0, 4, d, bx, kblk, &a0, dtype);
    if (rc) return rc;
  }
  if (C1 > 0) {
    long long d[4] = {C1, W, H, N};
    int bx[4] = {kblk, TILE_W, TILE_H, 1};
    int rc = get_map(src1, 4, d, bx, kblk, &a1, dtype);
    if (rc) return rc;
  } else {
    a1 = a0;
  }
  {
    long long d[3] = {CinP, CoutP, T};
    int bx[3] = {kblk, nt, 1};
    int rc = get_map(wpk_bf16, 3, d, bx, kblk, &b, dtype);
    if (rc) return rc;
  }
  ConvTcParams p;
  p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutP = CoutP; p.CoutStore = CoutStore; p.taps = T; p.ks = ksize;
  p.tiles_x = W / TILE_W; p.tiles_y = H / TILE_H; p.out_mode = out_mode; p.dt = dtype; p.dil = dilation; p.bias = bias; p.out_scale = nullptr; p.out = out;
  switch (kblk) {
    case 64: return dispatch_nt<64>(a0, a1, b, p, nt, stream);
    case 32: return dispatch_nt<32>(a0, a1, b, p, nt, stream);
    default: return dispatch_nt<16>(a0, a1, b, p, nt, stream);
  }
}

WSL_API int wsl_conv_tc(const void* src0, int C0, const void* src1, int C1, const void* wpk_bf16, const float* bias, void* out,
                        int out_mode, int N, int H, int W, int CoutP, int CoutStore, int ksize, int dtype, cudaStream_t stream) {
  return conv_tc_impl(src0, C0, src1, C1, wpk_bf16, bias, out, out_mode, N, H, W, CoutP, CoutStore, ksize, dtype, 1, stream);
}

// dilated 3x3 convolution (PNet2D, networks/pnet.py:25-28): same kernel, tap coordinates scaled by the dilation (TMA zero fill = padding)
WSL_API int wsl_conv_tc_dil(const void* src0, int C0, const void* src1, int C1, const void* wpk_bf16, const float* bias, void* out,
                            int out_mode, int N, int H, int W, int CoutP, int CoutStore, int ksize, int dtype, int dilation,
                            cudaStream_t stream) {
  return conv_tc_impl(src0, C0, src1, C1, wpk_bf16, bias, out, out_mode, N, H, W, CoutP, CoutStore, ksize, dtype, dilation, stream);
}

// pitch_x / pitch_dy > 0: the operands are channel-range views of wider channels-last tensors (fp16 hi/lo split planes)
static int wgrad_tc_impl(const void* src0, int C0, const void* src1, int C1, const void* dy, int CoutP, float* dw, int N, int H,
                         int W, int CoutReal, int ksize, int dtype, int pitch_x, int pitch_dy, const float* scale_a,
                         const float* scale_b, cudaStream_t stream, float* partial_ws = nullptr, long long partial_floats = 0,
                         int dilation = 1) {
  g_wgrad1_ws = partial_ws;
  g_wgrad1_ws_floats = partial_floats;
  WSL_REQUIRE(ksize == 3 || ksize == 1, "wsl_wgrad_tc: ksize must be 1 or 3");
  WSL_REQUIRE(dtype == 0 || dtype == 2, "wsl_wgrad_tc: dtype must be 0 (bf16) or 2 (fp16)");
  WSL_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "wsl_wgrad_tc: source channels must be multiples of 16 (got %d,%d)", C0, C1);
  WSL_REQUIRE(CoutP % 16 == 0, "wsl_wgrad_tc: CoutP must be a multiple of 16");
  WSL_REQUIRE(W % TILE_W == 0 && H % TILE_H == 0, "wsl_wgrad_tc: H,W must be multiples of the 8x16 pixel tile (got %dx%d)", H, W);
  const int cwa = CoutP >= 64 ? 64 : CoutP;       // 16 / 32 / 64
  WSL_REQUIRE(cwa == 16 || cwa == 32 || cwa == 64, "wsl_wgrad_tc: unsupported CoutP %d", CoutP);
  WSL_REQUIRE(CoutP < 128 || CoutP % 64 == 0, "wsl_wgrad_tc: CoutP >= 128 must be a multiple of 64");   // a half-filled last tile reads zeros (TMA)
  int cwb = 64;
  while (cwb > 16 && (C0 % cwb != 0 || (C1 > 0 && C1 % cwb != 0))) cwb >>= 1;
  const int na = CoutP >= 128 ? 2 : 1;
  const int m_tiles = CoutP >= 128 ? (CoutP + 127) / 128 : 1;
  CUtensorMap mdy, mx0, mx1;
  {
    long long d[4] = {CoutP, W, H, N};
    int bx[4] = {cwa, TILE_W, TILE_H, 1};
    int rc = get_map(dy, 4, d, bx, cwa, &mdy, dtype, pitch_dy);
    if (rc) return rc;
  }
  {
    long long d[4] = {C0, W, H, N};
    int bx[4] = {cwb, TILE_W, TILE_H, 1};
    int rc = get_map(src0, 4, d, bx, cwb, &mx0, dtype, pitch_x);
    if (rc) return rc;
  }
  if (C1 > 0) {
    long long d[4] = {C1, W, H, N};
    int bx[4] = {cwb, TILE_W, TILE_H, 1};
    int rc = get_map(src1, 4, d, bx, cwb, &mx1, dtype, pitch_x);
    if (rc) return rc;
  } else {
    mx1 = mx0;
  }
  WgradTcParams p;
  p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutP = CoutP; p.CoutReal = CoutReal; p.taps = ksize * ksize; p.ks = ksize;
  p.tiles_x = W / TILE_W; p.tiles_y = H / TILE_H; p.nchunks = N * p.tiles_x * p.tiles_y;
  p.n_tiles0 = C0 / cwb; p.n_tiles1 = C1 / cwb; p.tap_groups = ksize == 3 ? 3 : 1; p.dt = dtype; p.scale_a = scale_a; p.scale_b = scale_b; p.dw = dw; p.partials = nullptr; p.dil = dilation;
  if (cwa == 64 && na == 2) return wgrad_dispatch_b<64, 2>(cwb, mdy, mx0, mx1, p, m_tiles, stream);
  if (cwa == 64) return wgrad_dispatch_b<64, 1>(cwb, mdy, mx0, mx1, p, m_tiles, stream);
  if (cwa == 32) return wgrad_dispatch_b<32, 1>(cwb, mdy, mx0, mx1, p, m_tiles, stream);
  return wgrad_dispatch_b<16, 1>(cwb, mdy, mx0, mx1, p, m_tiles, stream);
}

WSL_API int wsl_wgrad_tc(const void* src0, int C0, const void* src1, int C1, const void* dy, int CoutP, float* dw, int N, int H,
                         int W, int CoutReal, int ksize, int dtype, float* partial_ws, long long partial_floats, cudaStream_t stream) {
  return wgrad_tc_impl(src0, C0, src1, C1, dy, CoutP, dw, N, H, W, CoutReal, ksize, dtype, 0, 0, nullptr, nullptr, stream, partial_ws,
                       partial_floats);
}

WSL_API int wsl_wgrad_tc_dil(const void* src0, int C0, const void* src1, int C1, const void* dy, int CoutP, float* dw, int N, int H,
                             int W, int CoutReal, int ksize, int dtype, int dilation, float* partial_ws, long long partial_floats,
                             cudaStream_t stream) {
  WSL_REQUIRE(dilation >= 1 && dilation <= 64, "wsl_wgrad_tc_dil: dilation must be in 1..64 (got %d)", dilation);
  return wgrad_tc_impl(src0, C0, src1, C1, dy, CoutP, dw, N, H, W, CoutReal, ksize, dtype, 0, 0, nullptr, nullptr, stream, partial_ws,
                       partial_floats, dilation);
}

// ---- fp16 hi/lo split ("fp16x3") tensor-core parity mode ------------------------------------------------------------
// An fp32 value v is carried as hi = fp16(v), lo = fp16(v - hi) (22 significant bits); a product a*b is evaluated as
// a_hi*b_hi + a_lo*b_hi + a_hi*b_lo on kind::f16 with fp32 accumulation (the dropped lo*lo term is ~2^-22 relative).
// Staged operands are channels-last fp16 [P][2C] = (hi plane | lo plane), produced by wsl_split_f32 (net_ops.cu).
//   forward / data gradient: K = 3*Cin: source 0 = the whole staged tensor (hi | lo) against (w_hi | w_hi), source 1 = its hi
//   plane again (channel-range view, pitch 2*Cin) against w_lo; weights fp16 [T][CoutP][3*Cin] from wsl_pack_split_weights.
//   weight gradient: three accumulating launches (dy_hi,x_hi), (dy_hi,x_lo), (dy_lo,x_hi).
// out_mode 2: fp32 NHWC [P][CoutStore]; 1: fp32 NCHW.
WSL_API int wsl_conv_tc_split(const void* staged, int Cin, const float* inv_scale, const void* wpk3, const float* bias, float* out,
                              int out_mode, int N, int H, int W, int CoutP, int CoutStore, int ksize, int dilation, cudaStream_t stream) {
  WSL_REQUIRE(dilation >= 1 && dilation <= 64, "wsl_conv_tc_split: dilation must be in 1..64 (got %d)", dilation);
  WSL_REQUIRE(ksize == 3 || ksize == 1, "wsl_conv_tc_split: ksize must be 1 or 3");
  WSL_REQUIRE(Cin % 16 == 0 && Cin > 0 && CoutP % 16 == 0, "wsl_conv_tc_split: channel counts must be multiples of 16 (got %d,%d)", Cin, CoutP);
  WSL_REQUIRE(W % TILE_W == 0 && H % TILE_H == 0, "wsl_conv_tc_split: H,W must be multiples of the 8x16 pixel tile (got %dx%d)", H, W);
  WSL_REQUIRE(out_mode == 1 || out_mode == 2, "wsl_conv_tc_split: out_mode must be 1 (fp32 NCHW) or 2 (fp32 NHWC)");
  int kblk = 64;
  while (kblk > 16 && Cin % kblk != 0) kblk >>= 1;
  int nt = CoutP >= 128 ? 128 : CoutP;
  while (nt > 16 && CoutP % nt != 0) nt >>= 1;
  WSL_REQUIRE(CoutP % nt == 0 && (nt == 16 || nt == 32 || nt == 64 || nt == 128), "wsl_conv_tc_split: unsupported CoutP %d", CoutP);
  const int T = ksize * ksize;
  CUtensorMap a0, a1, b;
  {
    long long d[4] = {2LL * Cin, W, H, N};
    int bx[4] = {kblk, TILE_W, TILE_H, 1};
    int rc = get_map(staged, 4, d, bx, kblk, &a0, 2);
    if (rc) return rc;
  }
  {
    long long d[4] = {Cin, W, H, N};
    int bx[4] = {kblk, TILE_W, TILE_H, 1};
    int rc = get_map(staged, 4, d, bx, kblk, &a1, 2, 2 * Cin);
    if (rc) return rc;
  }
  {
    long long d[3] = {3LL * Cin, CoutP, T};
    int bx[3] = {kblk, nt, 1};
    int rc = get_map(wpk3, 3, d, bx, kblk, &b, 2);
    if (rc) return rc;
  }
  ConvTcParams p;
  p.N = N; p.H = H; p.W = W; p.C0 = 2 * Cin; p.C1 = Cin; p.CoutP = CoutP; p.CoutStore = CoutStore; p.taps = T; p.ks = ksize;
  p.tiles_x = W / TILE_W; p.tiles_y = H / TILE_H; p.out_mode = out_mode; p.dt = 2; p.dil = dilation; p.bias = bias; p.out_scale = inv_scale; p.out = out;
  switch (kblk) {
    case 64: return dispatch_nt<64>(a0, a1, b, p, nt, stream);
    case 32: return dispatch_nt<32>(a0, a1, b, p, nt, stream);
    default: return dispatch_nt<16>(a0, a1, b, p, nt, stream);
  }
}

WSL_API int wsl_wgrad_tc_split(const void* x_staged, int Cin, const float* x_inv_scale, const void* dy_staged, int CoutP,
                               const float* dy_inv_scale, float* dw, int N, int H, int W, int CoutReal, int ksize, int dilation,
                               cudaStream_t stream) {
  WSL_REQUIRE(Cin % 16 == 0 && Cin > 0, "wsl_wgrad_tc_split: Cin must be a multiple of 16 (got %d)", Cin);
  const __half* x = reinterpret_cast<const __half*>(x_staged);
  const __half* g = reinterpret_cast<const __half*>(dy_staged);
  int rc = wgrad_tc_impl(x, Cin, nullptr, 0, g, CoutP, dw, N, H, W, CoutReal, ksize, 2, 2 * Cin, 2 * CoutP, dy_inv_scale, x_inv_scale, stream, nullptr, 0, dilation);
  if (rc) return rc;
  rc = wgrad_tc_impl(x + Cin, Cin, nullptr, 0, g, CoutP, dw, N, H, W, CoutReal, ksize, 2, 2 * Cin, 2 * CoutP, dy_inv_scale, x_inv_scale, stream, nullptr, 0, dilation);
  if (rc) return rc;
  return wgrad_tc_impl(x, Cin, nullptr, 0, g + CoutP, CoutP, dw, N, H, W, CoutReal, ksize, 2, 2 * Cin, 2 * CoutP, dy_inv_scale, x_inv_scale, stream, nullptr, 0, dilation);
}

WSL_API int wsl_channel_sum(const void* x, int dtype, long long P, int C, int Creal, float* out, float* ws, cudaStream_t stream) {
  WSL_REQUIRE(C % 8 == 0 && C / 8 <= 256 && Creal <= C, "wsl_channel_sum: unsupported channel count %d", C);
  const int rows = 256 / (C / 8);
  long long b = (P + rows * 16 - 1) / (rows * 16);
  if (b > 148 * 2) b = 148 * 2;
  if (b < 1) b = 1;
  if (dtype == 1) channel_sum_kernel<float><<<(int)b, 256, 256 * 8 * sizeof(float), stream>>>((const float*)x, P, C, Creal, out, ws);
  else if (dtype == 2) channel_sum_kernel<__nv_bfloat16, true><<<(int)b, 256, 256 * 8 * sizeof(float), stream>>>((const __nv_bfloat16*)x, P, C, Creal, out, ws);
  else channel_sum_kernel<__nv_bfloat16><<<(int)b, 256, 256 * 8 * sizeof(float), stream>>>((const __nv_bfloat16*)x, P, C, Creal, out, ws);
  return wsl_check_launch("channel_sum");
}

// conv_tc v2 entry point (same contract as wsl_conv_tc; needs W % 8 == 0 and H % (16*MT) == 0)
WSL_API int wsl_conv_tc2(const void* src0, int C0, const void* src1, int C1, const void* wpk_bf16, const float* bias, void* out,
                         int out_mode, int N, int H, int W, int CoutP, int CoutStore, int ksize, int dtype, float* stat_partials,
                         int* stat_rows_host, cudaStream_t stream) {
  WSL_REQUIRE(ksize == 3 || ksize == 1, "wsl_conv_tc2: ksize must be 1 or 3");
  WSL_REQUIRE(dtype == 0 || dtype == 2, "wsl_conv_tc2: dtype must be 0 (bf16) or 2 (fp16)");
  WSL_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "wsl_conv_tc2: source channels must be multiples of 16 (got %d,%d)", C0, C1);
  WSL_REQUIRE(CoutP % 16 == 0, "wsl_conv_tc2: CoutP must be a multiple of 16");
  {   // narrow outputs at >= 128-pixel rows: the row kernel (filter rows in N) is 1.8-2.3x faster per pixel
    static const bool row_on = [] { const char* e = getenv("WSL4MIS_CONV_ROW"); return !(e && e[0] == '0'); }();
    static const bool wide_on = [] { const char* e = getenv("WSL4MIS_ROW_KBW32"); return !(e && e[0] == '0'); }();
    // channel block per TMA box / swizzle span: 32 channels (64 B) when every source allows it and the output is 32 wide
    const int kbw = (wide_on && CoutP == 32 && C0 % 32 == 0 && C1 % 32 == 0) ? 32 : 16;
    const int KB = (C0 + C1) / kbw;
    // row segments of 32 output rows (6 % halo re-reads) unless that leaves fewer than ~4 work items per SM
    const int RS = (H % 32 == 0 && (long long)N * (W / ROW_PX) * (H / 32) >= 4 * 148) ? 32 : 16;
    if (row_on && ksize == 3 && (CoutP == 16 || CoutP == 32) && W % ROW_PX == 0 && H % 16 == 0 && (C0 + C1) <= 64) {
      CUtensorMap a0, a1, b;
      {
        long long d[4] = {C0, W, H, N};
        int bx[4] = {kbw, 130, 1, 1};
        int rc = get_map(src0, 4, d, bx, kbw, &a0, dtype);
        if (rc) return rc;
      }
      if (C1 > 0) {
        long long d[4] = {C1, W, H, N};
        int bx[4] = {kbw, 130, 1, 1};
        int rc = get_map(src1, 4, d, bx, kbw, &a1, dtype);
        if (rc) return rc;
      } else {
        a1 = a0;
      }
      {
        long long d[3] = {C0 + C1, CoutP, 9};
        int bx[3] = {kbw, CoutP, 1};
        int rc = get_map(wpk_bf16, 3, d, bx, kbw, &b, dtype);
        if (rc) return rc;
      }
      ConvV2Params p;
      p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutP = CoutP; p.CoutStore = CoutStore;
      p.tiles_x = 0; p.tiles_y = 0; p.ntiles = 0;
      p.out_mode = out_mode; p.desc_mode = 0; p.dt = dtype; p.bias = bias; p.out = out;
      p.stat_partials = (out_mode == 0) ? stat_partials : nullptr;
      const int rc = (CoutP == 16) ? launch_conv_row<16>(a0, a1, b, p, KB, RS, kbw, stream)
                                   : launch_conv_row<32>(a0, a1, b, p, KB, RS, kbw, stream);
      if (stat_rows_host) *stat_rows_host = g_conv2_last_rows;
      return rc;
    }
  }
  int kblk = 64;
  while (kblk > 16 && (C0 % kblk != 0 || (C1 > 0 && C1 % kblk != 0))) kblk >>= 1;
  int mt = kblk == 64 ? 1 : kblk == 32 ? 2 : 4;
  while (mt > 1 && H % (16 * mt) != 0) mt >>= 1;
  WSL_REQUIRE(W % 8 == 0 && H % (16 * mt) == 0, "wsl_conv_tc2: H,W must be multiples of the 16x8 pixel tile (got %dx%d)", H, W);
  const int CinP = C0 + C1, T = ksize * ksize;
  // N tile: largest of 128/64/32/16 dividing CoutP whose resident weight slice fits (~148 KB) and TMEM (2*MT*NT <= 512)
  int nt = CoutP >= 128 ? 128 : CoutP;
  while (nt > 16 && (CoutP % nt != 0 || (long long)T * CinP * nt * 2 > 148 * 1024 || 2 * mt * nt > 512 || (nt == 128 && mt > 2))) nt >>= 1;
  WSL_REQUIRE(CoutP % nt == 0 && (long long)T * CinP * nt * 2 <= 152 * 1024, "wsl_conv_tc2: weights do not fit (Cin %d, NT %d)", CinP, nt);
  // base_offset stays 0: the swizzle XOR is taken from absolute shared-memory address bits by both TMA and tcgen05
  // (verified on B200: setting base_offset = (addr>>7)&7 for the row-shifted views breaks every case).
  const int desc_mode = 0;
  const int pad = ksize / 2;
  CUtensorMap a0, a1, b;
  {
    long long d[4] = {C0, W, H, N};
    int bx[4] = {kblk, 8 + 2 * pad, 16 * mt + 2 * pad, 1};
    int rc = get_map(src0, 4, d, bx, kblk, &a0, dtype);
    if (rc) return rc;
  }
  if (C1 > 0) {
    long long d[4] = {C1, W, H, N};
    int bx[4] = {kblk, 8 + 2 * pad, 16 * mt + 2 * pad, 1};
    int rc = get_map(src1, 4, d, bx, kblk, &a1, dtype);
    if (rc) return rc;
  } else {
    a1 = a0;
  }
  {
    long long d[3] = {CinP, CoutP, T};
    int bx[3] = {kblk, nt, 1};
    int rc = get_map(wpk_bf16, 3, d, bx, kblk, &b, dtype);
    if (rc) return rc;
  }
  ConvV2Params p;
  p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutP = CoutP; p.CoutStore = CoutStore;
  p.tiles_x = W / 8; p.tiles_y = H / (16 * mt); p.ntiles = N * p.tiles_x * p.tiles_y;
  p.out_mode = out_mode; p.desc_mode = desc_mode; p.dt = dtype; p.bias = bias; p.out = out;
  p.stat_partials = (out_mode == 0) ? stat_partials : nullptr;
#define WSL_C2(KS_, KB_, MT_)                                                   \
  do {                                                                         \
    int rc_ = conv2_dispatch_nt<KS_, KB_, MT_>(nt, a0, a1, b, p, stream);      \
    if (stat_rows_host) *stat_rows_host = g_conv2_last_rows;                   \
    return rc_;                                                                \
  } while (0)
  if (ksize == 3) {
    if (kblk == 64) WSL_C2(3, 64, 1);
    if (kblk == 32) { if (mt == 2) WSL_C2(3, 32, 2); WSL_C2(3, 32, 1); }
    if (mt == 4) WSL_C2(3, 16, 4);
    if (mt == 2) WSL_C2(3, 16, 2);
    WSL_C2(3, 16, 1);
  } else {
    if (kblk == 64) WSL_C2(1, 64, 1);
    if (kblk == 32) { if (mt == 2) WSL_C2(1, 32, 2); WSL_C2(1, 32, 1); }
    if (mt == 4) WSL_C2(1, 16, 4);
    if (mt == 2) WSL_C2(1, 16, 2);
    WSL_C2(1, 16, 1);
  }
#undef WSL_C2
}

// 3x3 weight gradient, v2 (halo views).  Same contract as wsl_wgrad_tc; needs W % 8 == 0 and H % 16 == 0.
WSL_API int wsl_wgrad_tc2(const void* src0, int C0, const void* src1, int C1, const void* dy, int CoutP, float* dw, int N, int H,
                          int W, int CoutReal, int ksize, int dtype, cudaStream_t stream) {
  WSL_REQUIRE(ksize == 3, "wsl_wgrad_tc2: 3x3 only");
  WSL_REQUIRE(dtype == 0 || dtype == 2, "wsl_wgrad_tc2: dtype must be 0 (bf16) or 2 (fp16)");
  WSL_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "wsl_wgrad_tc2: source channels must be multiples of 16 (got %d,%d)", C0, C1);
  WSL_REQUIRE(CoutP % 16 == 0 && (CoutP < 128 || CoutP % 128 == 0), "wsl_wgrad_tc2: unsupported CoutP %d", CoutP);
  WSL_REQUIRE(W % 8 == 0 && H % 16 == 0, "wsl_wgrad_tc2: H,W must be multiples of the 16x8 pixel chunk (got %dx%d)", H, W);
  const int cwa = CoutP >= 64 ? 64 : CoutP;
  const int cwb = (C0 % 32 == 0 && (C1 == 0 || C1 % 32 == 0)) ? 32 : 16;
  const int na = CoutP >= 128 ? 2 : 1;
  const int m_tiles = CoutP >= 128 ? CoutP / 128 : 1;
  CUtensorMap mdy, mx0, mx1;
  {
    long long d[4] = {CoutP, W, H, N};
    int bx[4] = {cwa, 8, 16, 1};
    int rc = get_map(dy, 4, d, bx, cwa, &mdy, dtype);
    if (rc) return rc;
  }
  {
    long long d[4] = {C0, W, H, N};
    int bx[4] = {cwb, 10, 18, 1};
    int rc = get_map(src0, 4, d, bx, cwb, &mx0, dtype);
    if (rc) return rc;
  }
  if (C1 > 0) {
    long long d[4] = {C1, W, H, N};
    int bx[4] = {cwb, 10, 18, 1};
    int rc = get_map(src1, 4, d, bx, cwb, &mx1, dtype);
    if (rc) return rc;
  } else {
    mx1 = mx0;
  }
  WgradTcParams p;
  p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutP = CoutP; p.CoutReal = CoutReal; p.taps = 9; p.ks = 3;
  p.tiles_x = W / 8; p.tiles_y = H / 16; p.nchunks = N * p.tiles_x * p.tiles_y;
  p.n_tiles0 = C0 / cwb; p.n_tiles1 = C1 / cwb; p.tap_groups = 1; p.dt = dtype; p.scale_a = nullptr; p.scale_b = nullptr; p.dw = dw; p.partials = nullptr; p.dil = 1;
  if (cwb == 32) {
    if (cwa == 64 && na == 2) return launch_wgrad2<64, 2, 32>(mdy, mx0, mx1, p, m_tiles, stream);
    if (cwa == 64) return launch_wgrad2<64, 1, 32>(mdy, mx0, mx1, p, m_tiles, stream);
    if (cwa == 32) return launch_wgrad2<32, 1, 32>(mdy, mx0, mx1, p, m_tiles, stream);
    return launch_wgrad2<16, 1, 32>(mdy, mx0, mx1, p, m_tiles, stream);
  }
  if (cwa == 64 && na == 2) return launch_wgrad2<64, 2, 16>(mdy, mx0, mx1, p, m_tiles, stream);
  if (cwa == 64) return launch_wgrad2<64, 1, 16>(mdy, mx0, mx1, p, m_tiles, stream);
  if (cwa == 32) return launch_wgrad2<32, 1, 16>(mdy, mx0, mx1, p, m_tiles, stream);
  return launch_wgrad2<16, 1, 16>(mdy, mx0, mx1, p, m_tiles, stream);
}

// 3x3 weight gradient, v3 (filter columns in the MMA's M dimension).  Same contract as wsl_wgrad_tc2.
WSL_API int wsl_wgrad_tc3(const void* src0, int C0, const void* src1, int C1, const void* dy, int CoutP, float* dw, int N, int H,
                          int W, int CoutReal, int ksize, int dtype, float* partial_ws, long long partial_floats, cudaStream_t stream) {
  WSL_REQUIRE(ksize == 3, "wsl_wgrad_tc3: 3x3 only");
  WSL_REQUIRE(dtype == 0 || dtype == 2, "wsl_wgrad_tc3: dtype must be 0 (bf16) or 2 (fp16)");
  WSL_REQUIRE(C0 % 16 == 0 && C1 % 16 == 0 && C0 > 0, "wsl_wgrad_tc3: source channels must be multiples of 16 (got %d,%d)", C0, C1);
  WSL_REQUIRE(CoutP % 16 == 0 && (CoutP <= 64 || CoutP % 128 == 0), "wsl_wgrad_tc3: unsupported CoutP %d", CoutP);
  WSL_REQUIRE(CoutP == 16 || CoutP == 32 || CoutP >= 64, "wsl_wgrad_tc3: unsupported CoutP %d", CoutP);
  WSL_REQUIRE(W % 8 == 0 && H % 16 == 0, "wsl_wgrad_tc3: H,W must be multiples of the 16x8 pixel chunk (got %dx%d)", H, W);
  const int cwb = (C0 % 32 == 0 && (C1 == 0 || C1 % 32 == 0)) ? 32 : 16;
  const int cwn = CoutP >= 64 ? 64 : CoutP;
  const int ntile = CoutP >= 128 ? 128 : CoutP;
  const int n_tiles = CoutP / ntile;
  const bool dyn = ntile < 128;          // single dY box: filter rows merged into the MMA's N dimension (18-row dY box, 16-row X box)
  CUtensorMap mdy, mx0, mx1;
  {
    long long d[4] = {CoutP, W, H, N};
    int bx[4] = {cwn, 8, dyn ? 18 : 16, 1};
    int rc = get_map(dy, 4, d, bx, cwn, &mdy, dtype);
    if (rc) return rc;
  }
  {
    long long d[4] = {C0, W, H, N};
    int bx[4] = {cwb, 10, dyn ? 16 : 18, 1};
    int rc = get_map(src0, 4, d, bx, cwb, &mx0, dtype);
    if (rc) return rc;
  }
  if (C1 > 0) {
    long long d[4] = {C1, W, H, N};
    int bx[4] = {cwb, 10, dyn ? 16 : 18, 1};
    int rc = get_map(src1, 4, d, bx, cwb, &mx1, dtype);
    if (rc) return rc;
  } else {
    mx1 = mx0;
  }
  Wgrad3Params p;
  p.N = N; p.H = H; p.W = W; p.C0 = C0; p.C1 = C1; p.CoutReal = CoutReal;
  p.tiles_x = W / 8; p.tiles_y = H / 16; p.nchunks = N * p.tiles_x * p.tiles_y;
  p.k_tiles0 = C0 / cwb; p.k_tiles1 = C1 / cwb; p.dt = dtype; p.dw = dw; p.partials = nullptr;
#define WSL_W3(CWB_) \
  if (ntile == 128) return launch_wgrad3<CWB_, 64, 2>(mdy, mx0, mx1, p, n_tiles, partial_ws, partial_floats, stream); \
  if (ntile == 64) return launch_wgrad3<CWB_, 64, 1>(mdy, mx0, mx1, p, n_tiles, partial_ws, partial_floats, stream);  \
  if (ntile == 32) return launch_wgrad3<CWB_, 32, 1>(mdy, mx0, mx1, p, n_tiles, partial_ws, partial_floats, stream);  \
  return launch_wgrad3<CWB_, 16, 1>(mdy, mx0, mx1, p, n_tiles, partial_ws, partial_floats, stream);
  if (cwb == 32) { WSL_W3(32) }
  WSL_W3(16)
#undef WSL_W3
}

