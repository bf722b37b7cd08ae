// tcgen05 implicit-GEMM 3x3/1x1 convolution, second generation: operand reuse in shared memory.
//
// Same contract as conv_tc.cu (multi-segment K, fused bias/temb/residual/scale epilogue, GroupNorm partials),
// but built around two measurements of the v1 kernel (profiles/r01_conv_tc_v1_ncu.txt): it needs 128 B of
// L2->SMEM traffic per MMA clock per SM, and its epilogue (residual round trip through smem, per-column
// global loads of bias/temb with a 4 KB L1, serialized phases) is as long as the main loop.
//
//   * CTA tile = 8 (W) x 16*SUBS (H) pixels of one utterance = SUBS accumulators of 128 pixels x 128 channels.
//   * Per (64-channel chunk, dx in {-1,0,1}) ONE TMA box {64 ch, 8 px, 16*SUBS+2 rows} is loaded.  Because a pixel
//     row of the box is exactly 8 x 128 B = one 1024-B swizzle atom, the three dy taps (and the SUBS sub-tiles) are
//     the same smem bytes viewed through UMMA descriptors whose start address is advanced by whole atoms
//     ((sub*16 + dy + 1) * 1024 B) -- activation traffic drops from 9 to 3*(16*SUBS+2)/(16*SUBS) tile loads.
//   * Each weight tile (128 cout x 64 cin of one tap) feeds SUBS accumulators.
//     => L2->SMEM bytes per MMA clock: 128 (v1) -> 53 (SUBS=2).
//   * Separate smem rings for activations (A_STAGES x 34 KB) and weights (B_STAGES x 16 KB).
//   * The residual `x` of `(x + conv(h)) / sqrt(2)` enters as one more 1x1 K segment against an identity block
//     appended to the packed weights (exact: x * 1.0 accumulated in fp32), so the epilogue never reads it.
//   * Epilogue works in 64-channel halves that ping-pong between two 16 KB staging buffers: tcgen05.ld (64 columns)
//     -> + (bias + temb) from smem -> * scale -> fp16 -> swizzled smem -> TMA store, while the previous half's store
//     drains; GroupNorm partials are reduced from the staged half.
//
// TMEM: 2 stages x SUBS accumulators x 128 fp32 columns (= all 512 columns for SUBS=2).
#include "kernels.h"

namespace sgmse {

CUtensorMap make_act_map(const void* p, int N, int H, int W, int C, int bw, int bh, int bn);
CUtensorMap make_w_map(const void* p, int Cout, int Ktot, int block_n);
int num_sms();

namespace {

constexpr int BLOCK_N = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;
constexpr int ROW_BYTES = 8 * 128;                 // one pixel row of the box: 8 px x 64 ch fp16 = one swizzle atom
constexpr int SUB_BYTES = 16 * ROW_BYTES;          // 128 pixels x 64 channels (16 KB)
constexpr int B_BYTES = BLOCK_N * 128;             // 16 KB
constexpr int MAX_SEG = 4;

struct Tc2Params {
  int tiles_w, tiles_h;          // per utterance
  int num_m_tiles, num_tiles, n_tiles_n;
  int N, Cout;
  int nseg;
  int seg_chunks[MAX_SEG];
  int seg_taps[MAX_SEG];
  int seg_kb0[MAX_SEG];          // first K block of the segment
  const float* bias;
  const float* temb;
  int temb_stride;
  float scale;
  float* stats;
  int slots;
  int* dbg;
};

template <int SUBS, int A_STAGES, int B_STAGES>
struct Smem2 {
  static constexpr int A_ROWS = 16 * SUBS + 2;
  static constexpr int A_BYTES = A_ROWS * ROW_BYTES;
  static constexpr int OFF_B = A_STAGES * A_BYTES;
  static constexpr int OFF_STAGING = OFF_B + B_STAGES * B_BYTES;         // 2 x 16 KB (ping-pong halves)
  static constexpr int OFF_STATS = OFF_STAGING + 2 * SUB_BYTES;          // float [4][64][2]
  static constexpr int OFF_BIAS = OFF_STATS + 4 * 64 * 2 * 4;            // float [128]
  static constexpr int OFF_BARS = OFF_BIAS + BLOCK_N * 4;
  static constexpr int NUM_BARS = 2 * A_STAGES + 2 * B_STAGES + 4;
  static constexpr int OFF_TMEM_PTR = OFF_BARS + NUM_BARS * 8;
  static constexpr int TOTAL = OFF_TMEM_PTR + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static_assert(DYN_BYTES <= 232448, "shared memory budget exceeded");
};

__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, M128 N128

__device__ __forceinline__ void tmem_ld_32x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
        "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
        "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
        "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
        "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

template <int SUBS, int A_STAGES, int B_STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_d,
                const Tc2Params P) {
  using L = Smem2<SUBS, A_STAGES, B_STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BARS);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* b_full = a_empty + A_STAGES;
  uint64_t* b_empty = b_full + B_STAGES;
  uint64_t* tmem_full = b_empty + B_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
  uint8_t* staging = smem + L::OFF_STAGING;
  float* stats_sm = reinterpret_cast<float*>(smem + L::OFF_STATS);
  float* bias_sm = reinterpret_cast<float*>(smem + L::OFF_BIAS);
  constexpr uint32_t TMEM_COLS = 2 * SUBS * BLOCK_N;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_b); tma_prefetch_desc(&map_d);
    if (P.nseg > 1) tma_prefetch_desc(&map_a1);
    if (P.nseg > 2) tma_prefetch_desc(&map_a2);
    if (P.nseg > 3) tma_prefetch_desc(&map_a3);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_per_utt = P.tiles_w * P.tiles_h;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int m_tile = tile % P.num_m_tiles, n_tile = tile / P.num_m_tiles;
        const int n = m_tile / tiles_per_utt, rem = m_tile % tiles_per_utt;
        const int x0 = (rem % P.tiles_w) * 8, y0 = (rem / P.tiles_w) * (16 * SUBS);
        for (int s = 0; s < P.nseg; ++s) {
          const CUtensorMap* ma = s == 0 ? &map_a0 : (s == 1 ? &map_a1 : (s == 2 ? &map_a2 : &map_a3));
          const int nd = P.seg_taps[s] == 9 ? 3 : 1;
          const int chunks = P.seg_chunks[s];
          for (int ch = 0; ch < chunks; ++ch)
            for (int dxi = 0; dxi < nd; ++dxi) {
              const int dx = nd == 3 ? dxi - 1 : 0;
              mbar_wait(&a_empty[sa], pa ^ 1, P.dbg, 100 + sa);
              mbar_arrive_expect_tx(&a_full[sa], L::A_BYTES);
              tma_load_4d(smem + sa * L::A_BYTES, ma, &a_full[sa], ch * BLOCK_K, x0 + dx, y0 - 1, n);
              if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
              for (int dyi = 0; dyi < nd; ++dyi) {
                const int tap = nd == 3 ? dyi * 3 + dxi : 0;
                const int kb = P.seg_kb0[s] + tap * chunks + ch;
                mbar_wait(&b_empty[sb], pb ^ 1, P.dbg, 150 + sb);
                mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
                tma_load_2d(smem + L::OFF_B + sb * B_BYTES, &map_b, &b_full[sb], kb * BLOCK_K, n_tile * BLOCK_N);
                if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
              }
            }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    int sa = 0; uint32_t pa = 0;
    int sb = 0; uint32_t pb = 0;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], as_phase ^ 1, P.dbg, 200 + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * SUBS * BLOCK_N);
      uint32_t first = 1;
      for (int s = 0; s < P.nseg; ++s) {
        const int nd = P.seg_taps[s] == 9 ? 3 : 1;
        const int chunks = P.seg_chunks[s];
        const bool last_seg = s == P.nseg - 1;
        for (int ch = 0; ch < chunks; ++ch)
          for (int dxi = 0; dxi < nd; ++dxi) {
            mbar_wait(&a_full[sa], pa, P.dbg, 300 + sa);
            const uint32_t a_base = smem_u32(smem + sa * L::A_BYTES);
            for (int dyi = 0; dyi < nd; ++dyi) {
              mbar_wait(&b_full[sb], pb, P.dbg, 350 + sb);
              tc_fence_after();
              if (lane == 0) {
                const int row0 = nd == 3 ? dyi : 1;      // box starts one pixel row above the tile
                const uint64_t bdesc = smem_desc_sw128(smem_u32(smem + L::OFF_B + sb * B_BYTES));
#pragma unroll
                for (int sub = 0; sub < SUBS; ++sub) {
                  const uint64_t adesc = smem_desc_sw128(a_base + (uint32_t)((sub * 16 + row0) * ROW_BYTES));
#pragma unroll
                  for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                    tc_mma_f16(d_tmem + (uint32_t)(sub * BLOCK_N), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC,
                               (first && k == 0) ? 0u : 1u);
                }
                tc_commit(&b_empty[sb]);
                const bool last = last_seg && ch == chunks - 1 && dxi == nd - 1 && dyi == nd - 1;
                if (dyi == nd - 1) tc_commit(&a_empty[sa]);
                if (last) tc_commit(&tmem_full[as]);
              }
              first = 0;
              __syncwarp();
              if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
            }
            if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          }
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int e = threadIdx.x - 64;
    const int lg = warp & 3;
    const int row = lg * 32 + lane;
    const int sw = row & 7;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int m_tile = tile % P.num_m_tiles, n_tile = tile / P.num_m_tiles;
      const int n = m_tile / tiles_per_utt, rem = m_tile % tiles_per_utt;
      const int tx = rem % P.tiles_w, ty = rem / P.tiles_w;
      const int x0 = tx * 8;
      const int c_tile = n_tile * BLOCK_N;
      {
        // bias + time-embedding bias of this (utterance, channel tile): one value per thread into smem
        float bt = P.bias ? __ldg(P.bias + c_tile + e) : 0.f;
        if (P.temb) bt += __ldg(P.temb + (size_t)n * P.temb_stride + c_tile + e);
        named_bar_sync(1, NUM_EPI_THREADS);      // previous tile's readers of bias_sm are done
        bias_sm[e] = bt;
      }
      mbar_wait(&tmem_full[as], as_phase, P.dbg, 500 + as);
      tc_fence_after();

#pragma unroll 1
      for (int sub = 0; sub < SUBS; ++sub) {
        const int y0 = (ty * SUBS + sub) * 16;
        const int slot = (ty * SUBS + sub) * P.tiles_w + tx;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint8_t* buf = staging + half * SUB_BYTES;
          // buffer `half` was last used two stores ago: at most the other half's store may still be reading smem
          if (e == 0) tma_store_wait_read1();
          named_bar_sync(1, NUM_EPI_THREADS);    // ... also: bias_sm visible, stats readers of `buf` done
          uint32_t r[64];
          tmem_ld_32x64(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)((as * SUBS + sub) * BLOCK_N + half * 64), r);
          tmem_ld_wait();
          if (sub == SUBS - 1 && half == 1) {
            // accumulator stage fully drained -> hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
          }
          uint8_t* my_row = buf + row * 128;
          const float4* bs = reinterpret_cast<const float4*>(bias_sm + half * 64);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 b0 = bs[2 * g], b1 = bs[2 * g + 1];
            uint4 ov;
            __half2* oh = reinterpret_cast<__half2*>(&ov);
            oh[0] = __floats2half2_rn((__uint_as_float(r[g * 8 + 0]) + b0.x) * P.scale, (__uint_as_float(r[g * 8 + 1]) + b0.y) * P.scale);
            oh[1] = __floats2half2_rn((__uint_as_float(r[g * 8 + 2]) + b0.z) * P.scale, (__uint_as_float(r[g * 8 + 3]) + b0.w) * P.scale);
            oh[2] = __floats2half2_rn((__uint_as_float(r[g * 8 + 4]) + b1.x) * P.scale, (__uint_as_float(r[g * 8 + 5]) + b1.y) * P.scale);
            oh[3] = __floats2half2_rn((__uint_as_float(r[g * 8 + 6]) + b1.z) * P.scale, (__uint_as_float(r[g * 8 + 7]) + b1.w) * P.scale);
            *reinterpret_cast<uint4*>(my_row + ((g ^ sw) << 4)) = ov;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, NUM_EPI_THREADS);
          if (e == 0) {
            tma_store_4d(&map_d, buf, c_tile + half * 64, x0, y0, n);
            tma_store_commit();
          }
          if (P.stats) {
            // per-channel (sum, sum^2) over each 32-row segment of the staged half (the fp16 values actually stored)
            const int seg = e >> 5;
            float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              const int r2 = seg * 32 + rr;
              const __half2 h = *reinterpret_cast<const __half2*>(buf + r2 * 128 + ((((lane >> 2) ^ (r2 & 7)) << 4) | ((lane & 3) << 2)));
              const float2 f = __half22float2(h);
              s0 += f.x; q0 += f.x * f.x; s1 += f.y; q1 += f.y * f.y;
            }
            float* d = stats_sm + ((seg * 64) + lane * 2) * 2;
            d[0] = s0; d[1] = q0; d[2] = s1; d[3] = q1;
            named_bar_sync(1, NUM_EPI_THREADS);
            if (e < 64) {
              float s = 0.f, q = 0.f;
#pragma unroll
              for (int j = 0; j < 4; ++j) { s += stats_sm[((j * 64) + e) * 2]; q += stats_sm[((j * 64) + e) * 2 + 1]; }
              float* o = P.stats + (((size_t)n * P.slots + slot) * P.Cout + c_tile + half * 64 + e) * 2;
              o[0] = s; o[1] = q;
            }
          }
        }
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
    if (e == 0) tma_store_wait_all0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int SUBS, int A_STAGES, int B_STAGES>
void launch2(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  using L = Smem2<SUBS, A_STAGES, B_STAGES>;
  Tc2Params P{};
  P.tiles_w = out.W / 8; P.tiles_h = out.H / (16 * SUBS);
  P.num_m_tiles = P.tiles_w * P.tiles_h * out.N;
  P.n_tiles_n = out.C / BLOCK_N;
  P.num_tiles = P.num_m_tiles * P.n_tiles_n;
  P.N = out.N; P.Cout = out.C;
  // segments: the conv's own, plus the residual as a 1x1 segment against the identity tail of the weights
  const TensorDesc* srcs[MAX_SEG];
  int taps[MAX_SEG];
  int nseg = 0;
  for (int i = 0; i < a.nseg; ++i) { srcs[nseg] = &a.seg[i].src; taps[nseg++] = a.seg[i].taps; }
  if (a.residual) { srcs[nseg] = a.residual; taps[nseg++] = 1; }
  P.nseg = nseg;
  CUtensorMap ma[MAX_SEG];
  int kb = 0;
  for (int i = 0; i < MAX_SEG; ++i) {
    const TensorDesc& s = *srcs[i < nseg ? i : 0];
    ma[i] = make_act_map(s.p, s.N, s.H, s.W, s.C, 8, L::A_ROWS, 1);
    if (i < nseg) {
      P.seg_chunks[i] = s.C / 64; P.seg_taps[i] = taps[i]; P.seg_kb0[i] = kb;
      kb += taps[i] * (s.C / 64);
    }
  }
  const int ld = a.w_tc_ld ? a.w_tc_ld : a.ktot();
  SG_CHECK(kb * 64 <= ld, "conv_tc2: K blocks (%d) exceed the packed weight row (%d)", kb * 64, ld);
  const CUtensorMap mb = make_w_map(a.w_tc, out.C, ld, BLOCK_N);
  const CUtensorMap md = make_act_map(out.p, out.N, out.H, out.W, out.C, 8, 16, 1);
  P.bias = a.bias; P.temb = a.temb; P.temb_stride = a.temb_stride;
  P.scale = a.scale;
  out.slots = P.tiles_w * P.tiles_h * SUBS;
  P.stats = out.stats; P.slots = out.slots;
  P.dbg = dbg;
  auto kern = conv_tc2_kernel<SUBS, A_STAGES, B_STAGES>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs)) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
  }
  const int grid = P.num_tiles < num_sms() ? P.num_tiles : num_sms();
  kern<<<grid, NUM_THREADS, L::DYN_BYTES, st>>>(ma[0], ma[1], ma[2], ma[3], mb, md, P);
  CUDA_OK(cudaGetLastError());
}

}  // namespace

bool conv_tc2_supported(const ConvArgs& a, const TensorDesc& out) {
  if (out.dt != DT_F16 || out.C % 128 != 0 || a.w_tc == nullptr) return false;
  if (out.W % 8 != 0 || out.H % 16 != 0) return false;
  for (int i = 0; i < a.nseg; ++i)
    if (a.seg[i].src.C % 64 != 0 || a.seg[i].src.dt != DT_F16) return false;
  if (a.residual && (!a.tc_identity_tail || a.nseg >= MAX_SEG || a.residual->C != out.C)) return false;
  return true;
}

void launch_conv_tc2(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  SG_CHECK(conv_tc2_supported(a, out), "conv_tc2: unsupported shape");
  if (out.H % 32 == 0) launch2<2, 3, 5>(st, a, out, dbg);
  else launch2<1, 4, 6>(st, a, out, dbg);
}

}  // namespace sgmse
