// tcgen05 implicit-GEMM convolution for sm_100a (3x3 pad-1 / 1x1, NHWC fp16 -> fp16, fp32 accumulate in TMEM).
//
//   D[pixel, cout] = sum_{segment, tap, cin} A_seg[pixel + tap offset, cin] * W[cout, (segment, tap, cin)]
//
// GEMM view: M = N*H*W pixels (tiles of 128 = one TMA box bw x bh x bn of pixels), N = Cout (tiles of
// BLOCK_N), K = sum over segments of taps*Cin (blocks of 64 channels of one tap).  There is no im2col
// buffer: for each (tap, 64-channel chunk) the producer issues ONE 4-D TMA box load of the activation
// tensor at the tap-shifted pixel coordinates; TMA zero-fills out-of-bounds pixels, which is exactly
// the conv's zero padding, and lands the box as a 128-row K-major SWIZZLE_128B operand tile.  Several
// K segments let one accumulator take  conv3x3(h) + conv1x1(x_a) + conv1x1(x_b)  (ResnetBlockBigGANpp's
// Conv_1 plus the Conv_2 shortcut over a concatenated skip input) without materialising the concat.
//
// Persistent, warp-specialised CTA (192 threads, 1 CTA/SM):
//   warp 0      TMA producer          (smem ring: NUM_STAGES x {A 16 KB, B BLOCK_N*128 B})
//   warp 1      tcgen05.mma issuer    (one lane), TMEM alloc/dealloc; 2 accumulator stages in TMEM
//   warps 2..5  epilogue: tcgen05.ld -> +bias +time-embedding bias +residual, *scale -> fp16 ->
//               swizzled smem staging -> TMA store; per-channel (sum, sum^2) partials for the next
//               GroupNorm are reduced from the staged tile (deterministic, no atomics).
//
// Reference semantics: F.conv2d calls of ResnetBlockBigGANpp.forward (layerspp.py:260-269) and NIN
// (layers.py:546-555); the epilogue fuses `h += Dense_0(act(temb))[:, :, None, None]`, the Conv_2
// shortcut add and `(x + h) / sqrt(2)` (layerspp.py:262-274).
#include "kernels.h"

#include <mutex>

namespace sgmse {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;     // fp16 elements = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;
constexpr int A_TILE_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KB

struct TcParams {
  int tiles_w, tiles_h, tiles_n;
  int bw, bh, bn;
  int n_tiles_n, num_m_tiles, num_tiles;
  int N, Cout;
  int nseg;
  int seg_chunks[3];
  int seg_taps[3];
  int num_k_blocks;
  const float* bias;
  const float* temb;
  int temb_stride;
  int has_residual;
  float scale;
  float* stats;
  int slots;
  int* dbg;
};

template <int BLOCK_N, int NUM_STAGES>
struct SmemLayout {
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int N_CHUNKS = BLOCK_N / 64;
  static constexpr int STAGING_BYTES = N_CHUNKS * A_TILE_BYTES;
  static constexpr int OFF_STAGING = NUM_STAGES * STAGE_BYTES;
  static constexpr int OFF_STATS = OFF_STAGING + STAGING_BYTES;         // float [4][BLOCK_N][2]
  static constexpr int OFF_BARS = OFF_STATS + 4 * BLOCK_N * 2 * 4;
  static constexpr int NUM_BARS = 2 * NUM_STAGES + 5;
  static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;                        // slack for 1024 B alignment
};

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  // K-major, SWIZZLE_128B canonical layout: 8-row x 128 B atoms, atoms 1024 B apart along M/N.
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);      // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

template <int BLOCK_N>
__device__ __forceinline__ constexpr uint32_t make_idesc() {
  return (1u << 4)                    // D format: F32
         | (0u << 7) | (0u << 10)     // A, B format: F16
         | (0u << 15) | (0u << 16)    // A, B K-major
         | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

template <int BLOCK_N, int NUM_STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
               const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_d, const __grid_constant__ CUtensorMap map_r,
               const TcParams P) {
  using L = SmemLayout<BLOCK_N, NUM_STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BARS);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + NUM_STAGES;
  uint64_t* tmem_full = bars + 2 * NUM_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_full = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
  uint8_t* staging = smem + L::OFF_STAGING;
  float* stats_sm = reinterpret_cast<float*>(smem + L::OFF_STATS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_d);
    if (P.nseg > 1) tma_prefetch_desc(&map_a1);
    if (P.nseg > 2) tma_prefetch_desc(&map_a2);
    if (P.has_residual) tma_prefetch_desc(&map_r);
    for (int i = 0; i < NUM_STAGES; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_init(res_full, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, 2 * BLOCK_N);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                    // everything above is on-chip set-up; global memory is first touched below
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int m_tile = tile % P.num_m_tiles, n_tile = tile / P.num_m_tiles;
        const int tx = m_tile % P.tiles_w, ty = (m_tile / P.tiles_w) % P.tiles_h, tn = m_tile / (P.tiles_w * P.tiles_h);
        const int x0 = tx * P.bw, y0 = ty * P.bh, n0 = tn * P.bn;
        int kb = 0;
        for (int s = 0; s < P.nseg; ++s) {
          const CUtensorMap* ma = s == 0 ? &map_a0 : (s == 1 ? &map_a1 : &map_a2);
          const int taps = P.seg_taps[s];
          for (int tap = 0; tap < taps; ++tap) {
            const int dy = taps == 9 ? tap / 3 - 1 : 0, dx = taps == 9 ? tap % 3 - 1 : 0;
            for (int ch = 0; ch < P.seg_chunks[s]; ++ch, ++kb) {
              mbar_wait(&empty_bar[stage], phase ^ 1, P.dbg, 100 + stage);
              uint8_t* sa = smem + stage * L::STAGE_BYTES;
              mbar_arrive_expect_tx(&full_bar[stage], L::STAGE_BYTES);
              tma_load_4d(sa, ma, &full_bar[stage], ch * BLOCK_K, x0 + dx, y0 + dy, n0);
              tma_load_2d(sa + A_TILE_BYTES, &map_b, &full_bar[stage], kb * BLOCK_K, n_tile * BLOCK_N);
              if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc<BLOCK_N>();
    int stage = 0; uint32_t phase = 0;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], as_phase ^ 1, P.dbg, 200 + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * BLOCK_N);
      for (int kb = 0; kb < P.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase, P.dbg, 300 + stage);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + stage * L::STAGE_BYTES);
          const uint64_t adesc = make_smem_desc(a_addr);
          const uint64_t bdesc = make_smem_desc(a_addr + A_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 16 fp16 = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
            tc_mma_f16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          tc_commit(&empty_bar[stage]);                       // frees the smem stage when the MMAs retire
          if (kb == P.num_k_blocks - 1) tc_commit(&tmem_full[as]);  // accumulator complete
        }
        __syncwarp();
        if (++stage == NUM_STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int e = threadIdx.x - 64;             // 0..127
    const int lg = warp & 3;                    // TMEM lane group this warp may access
    const int row = lg * 32 + lane;             // accumulator row = pixel index inside the box
    const int rps = P.bw * P.bh;                // rows per sample inside a tile
    int as = 0; uint32_t as_phase = 0; uint32_t res_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int m_tile = tile % P.num_m_tiles, n_tile = tile / P.num_m_tiles;
      const int tx = m_tile % P.tiles_w, ty = (m_tile / P.tiles_w) % P.tiles_h, tn = m_tile / (P.tiles_w * P.tiles_h);
      const int x0 = tx * P.bw, y0 = ty * P.bh, n0 = tn * P.bn;
      const int c_tile = n_tile * BLOCK_N;

      // staging buffer must be free: the previous tile's TMA store has finished reading it
      if (e == 0) tma_store_wait_read0();
      named_bar_sync(1, NUM_EPI_THREADS);       // ... and every epilogue thread is done with the previous tile
      if (P.has_residual) {
        if (e == 0) {
          fence_proxy_async_smem();
          mbar_arrive_expect_tx(res_full, L::STAGING_BYTES);
#pragma unroll
          for (int c = 0; c < L::N_CHUNKS; ++c)
            tma_load_4d(staging + c * A_TILE_BYTES, &map_r, res_full, c_tile + c * 64, x0, y0, n0);
        }
        mbar_wait(res_full, res_phase, P.dbg, 400);
        res_phase ^= 1;
      }

      mbar_wait(&tmem_full[as], as_phase, P.dbg, 500 + as);
      tc_fence_after();

      int n_s = n0 + row / rps;
      if (n_s >= P.N) n_s = P.N - 1;            // rows of a partially out-of-range box (never stored)
      const float* temb_row = P.temb ? P.temb + (size_t)n_s * P.temb_stride + c_tile : nullptr;
      const float* bias_row = P.bias ? P.bias + c_tile : nullptr;
      const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * BLOCK_N);
      uint8_t* my_row = staging + row * 128;
      const int sw = row & 7;
#pragma unroll 1
      for (int c32 = 0; c32 < BLOCK_N / 32; ++c32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + (uint32_t)(c32 * 32), r);
        tmem_ld_wait();
        uint8_t* chunk_row = my_row + (c32 >> 1) * A_TILE_BYTES;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cb = c32 * 32 + g * 8;      // channel inside the tile
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float x = __uint_as_float(r[g * 8 + i]);
            if (bias_row) x += __ldg(bias_row + cb + i);
            if (temb_row) x += __ldg(temb_row + cb + i);
            v[i] = x;
          }
          uint4* sp = reinterpret_cast<uint4*>(chunk_row + ((((c32 & 1) * 4 + g) ^ sw) << 4));
          if (P.has_residual) {
            const uint4 rv = *sp;
            const __half2* rh = reinterpret_cast<const __half2*>(&rv);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(rh[i]); v[2 * i] += f.x; v[2 * i + 1] += f.y; }
          }
          uint4 ov;
          __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
          for (int i = 0; i < 4; ++i) oh[i] = __floats2half2_rn(v[2 * i] * P.scale, v[2 * i + 1] * P.scale);
          *sp = ov;
        }
      }
      // accumulator stage drained -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
      if (++as == 2) { as = 0; as_phase ^= 1; }

      fence_proxy_async_smem();
      named_bar_sync(1, NUM_EPI_THREADS);
      if (e == 0) {
#pragma unroll
        for (int c = 0; c < L::N_CHUNKS; ++c)
          tma_store_4d(&map_d, staging + c * A_TILE_BYTES, c_tile + c * 64, x0, y0, n0);
        tma_store_commit();
      }

      if (P.stats) {
        // per-channel (sum, sum^2) over each 32-row segment, from the fp16 values actually stored
        const int seg = e >> 5;
#pragma unroll
        for (int c = 0; c < L::N_CHUNKS; ++c) {
          float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
          const uint8_t* cbase = staging + c * A_TILE_BYTES;
#pragma unroll 8
          for (int rr = 0; rr < 32; ++rr) {
            const int r2 = seg * 32 + rr;
            const __half2 h = *reinterpret_cast<const __half2*>(cbase + r2 * 128 + ((((lane >> 2) ^ (r2 & 7)) << 4) | ((lane & 3) << 2)));
            const float2 f = __half22float2(h);
            s0 += f.x; q0 += f.x * f.x; s1 += f.y; q1 += f.y * f.y;
          }
          float* d = stats_sm + ((seg * BLOCK_N) + c * 64 + lane * 2) * 2;
          d[0] = s0; d[1] = q0; d[2] = s1; d[3] = q1;
        }
        named_bar_sync(1, NUM_EPI_THREADS);
        const int groups = rps >= 128 ? 1 : 128 / rps;       // samples covered by this tile
        const int spg = 4 / groups;                          // 32-row segments per sample
        const int slot = ty * P.tiles_w + tx;
        for (int c = e; c < BLOCK_N; c += NUM_EPI_THREADS) {
          for (int g = 0; g < groups; ++g) {
            const int n = n0 + g;
            if (n >= P.N) break;
            float s = 0.f, q = 0.f;
            for (int j = 0; j < spg; ++j) {
              s += stats_sm[(((g * spg + j) * BLOCK_N) + c) * 2];
              q += stats_sm[(((g * spg + j) * BLOCK_N) + c) * 2 + 1];
            }
            float* d = P.stats + (((size_t)n * P.slots + slot) * P.Cout + c_tile + c) * 2;
            d[0] = s; d[1] = q;
          }
        }
      }
    }
    if (e == 0) tma_store_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BLOCK_N);
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// host side (tensor-map helpers are shared with conv_tc2.cu)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  SG_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

// NHWC fp16 tensor, box {64 channels, bw, bh, bn}
CUtensorMap make_act_map(const void* p, int N, int H, int W, int C, int bw, int bh, int bn) {
  CUtensorMap m;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bn};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(p), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d box %dx%dx%d) failed: %d", N, H, W, C, bw, bh, bn, (int)r);
  return m;
}
// weights [Cout][Ktot] fp16, box {64, block_n}
CUtensorMap make_w_map(const void* p, int Cout, int Ktot, int block_n) {
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)Cout};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)block_n};
  cuuint32_t es[2] = {1, 1};
  CUresult r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(p), dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights %dx%d) failed: %d", Cout, Ktot, (int)r);
  return m;
}

// Pixel box (bw, bh, bn) with bw*bh*bn = 128, bw | W, bh | H, bn in {1,2,4}: prefer a single sample
// and a square-ish box (smallest halo when neighbouring tiles share L2 lines).
bool choose_box(int H, int W, int& bw, int& bh, int& bn) {
  int best = -1000000;
  bool found = false;
  for (int w = 1; w <= 128; w *= 2)
    for (int h = 1; w * h <= 128; h *= 2) {
      if (w * h < 32 || W % w || H % h) continue;
      const int n = 128 / (w * h);
      const int score = -n * 100 - (w > h ? w / h : h / w);
      if (score > best) { best = score; bw = w; bh = h; bn = n; found = true; }
    }
  return found;
}

int num_sms() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    CUDA_OK(cudaGetDevice(&dev));
    CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

namespace {

template <int BLOCK_N, int NUM_STAGES>
void launch_impl(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  using L = SmemLayout<BLOCK_N, NUM_STAGES>;
  int bw, bh, bn;
  SG_CHECK(choose_box(out.H, out.W, bw, bh, bn), "conv_tc: no 128-pixel box for %dx%d", out.H, out.W);
  TcParams P{};
  P.bw = bw; P.bh = bh; P.bn = bn;
  P.tiles_w = out.W / bw; P.tiles_h = out.H / bh; P.tiles_n = cdiv(out.N, bn);
  P.n_tiles_n = out.C / BLOCK_N;
  P.num_m_tiles = P.tiles_w * P.tiles_h * P.tiles_n;
  P.num_tiles = P.num_m_tiles * P.n_tiles_n;
  P.N = out.N; P.Cout = out.C;
  P.nseg = a.nseg;
  CUtensorMap ma[3];
  int kblocks = 0;
  for (int i = 0; i < 3; ++i) {
    const TensorDesc& s = a.seg[i < a.nseg ? i : 0].src;
    ma[i] = make_act_map(s.p, s.N, s.H, s.W, s.C, bw, bh, bn);
    if (i < a.nseg) {
      P.seg_chunks[i] = s.C / 64; P.seg_taps[i] = a.seg[i].taps;
      kblocks += a.seg[i].taps * (s.C / 64);
    }
  }
  P.num_k_blocks = kblocks;
  const CUtensorMap mb = make_w_map(a.w_tc, out.C, a.w_tc_ld ? a.w_tc_ld : a.ktot(), BLOCK_N);
  const CUtensorMap md = make_act_map(out.p, out.N, out.H, out.W, out.C, bw, bh, bn);
  const TensorDesc& rs = a.residual ? *a.residual : out;
  const CUtensorMap mr = make_act_map(rs.p, rs.N, rs.H, rs.W, rs.C, bw, bh, bn);
  P.bias = a.bias; P.temb = a.temb; P.temb_stride = a.temb_stride;
  P.has_residual = a.residual ? 1 : 0;
  P.scale = a.scale;
  out.slots = P.tiles_w * P.tiles_h;
  P.stats = out.stats; P.slots = out.slots;
  P.dbg = dbg;
  auto kern = conv_tc_kernel<BLOCK_N, NUM_STAGES>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs)) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
  }
  const int grid = P.num_tiles < num_sms() ? P.num_tiles : num_sms();
  launch_k(kern, dim3(grid), dim3(NUM_THREADS), (size_t)L::DYN_BYTES, st, ma[0], ma[1], ma[2], mb, md, mr, P);
  CUDA_OK(cudaGetLastError());
}

}  // namespace

thread_local int g_tc_variant = 0;   // see kernels.h
thread_local int g_tc1_narrow = 0;   // see launch_conv_tc
volatile int* g_wait_code_host = nullptr;

bool conv_tc_supported(const ConvArgs& a, const TensorDesc& out) {
  if (out.dt != DT_F16 || out.C % 64 != 0 || a.w_tc == nullptr) return false;
  for (int i = 0; i < a.nseg; ++i)
    if (a.seg[i].src.C % 64 != 0 || a.seg[i].src.dt != DT_F16) return false;
  int bw, bh, bn;
  return choose_box(out.H, out.W, bw, bh, bn);
}

void launch_conv_tc(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  SG_CHECK(conv_tc_supported(a, out), "conv_tc: unsupported shape (Cout=%d)", out.C);
  for (int i = 0; i < a.nseg; ++i) {
    const TensorDesc& s = a.seg[i].src;
    SG_CHECK(s.N == out.N && s.H == out.H && s.W == out.W, "conv_tc: segment %d shape mismatch", i);
  }
  if (a.residual) SG_CHECK(a.residual->C == out.C && a.residual->dt == DT_F16, "conv_tc: residual mismatch");
  if (a.gn_ab) {                                                      // the engine only asks for fusion when it applies
#ifdef SGMSE_B200_LAB        // the superseded generations (conv_tc2/3/5) live in the lab twin only (build.py)
    if (g_tc_variant == 5) { launch_conv_tc5(st, a, out, dbg); return; }
#endif
    launch_conv_tc6(st, a, out, dbg);
    return;
  }
  if ((g_tc_variant == 0 || g_tc_variant >= 6) && conv_tc6_supported(a, out)) { launch_conv_tc6(st, a, out, dbg); return; }
  if ((g_tc_variant == 0 || g_tc_variant >= 4) && conv_tc4_supported(a, out)) { launch_conv_tc4(st, a, out, dbg); return; }
#ifdef SGMSE_B200_LAB
  if (g_tc_variant == 3 && conv_tc3_supported(a, out)) { launch_conv_tc3(st, a, out, dbg); return; }
  if (g_tc_variant != 1 && conv_tc2_supported(a, out)) { launch_conv_tc2(st, a, out, dbg); return; }
#endif
  // tc1_narrow (round 2, on by default since; bit-identical, profiles/r02_candidates_gate.txt): on the levels below 16 rows a 128-wide channel tile leaves 32-64 CTAs that
  // each stream 32 KB per 64-deep k-block; 64-wide tiles double the CTAs and cut the per-CTA stream to 24 KB.  The
  // accumulation order of an output element does not depend on the tile width: bit-identical.
  if (out.C % 128 == 0 && g_tc1_narrow) {
    int bw, bh, bn;
    choose_box(out.H, out.W, bw, bh, bn);
    const int tiles128 = (out.W / bw) * (out.H / bh) * cdiv(out.N, bn) * (out.C / 128);
    if (tiles128 * 2 <= num_sms()) { launch_impl<64, 6>(st, a, out, dbg); return; }
  }
  if (out.C % 128 == 0) launch_impl<128, 5>(st, a, out, dbg);
  else launch_impl<64, 6>(st, a, out, dbg);
}

}  // namespace sgmse
