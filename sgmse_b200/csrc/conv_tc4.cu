// tcgen05 implicit-GEMM 3x3/1x1 convolution, fourth generation: operands swapped (D^T = W * A^T).
//
// What profiles/r01_conv_tc2_source.txt shows: in conv_tc2 the epilogue warps idle ~50 % of the time waiting for
// accumulators and the single MMA-issuing thread executes ~230 SASS instructions per K block (8 UMMAs of
// M128 x N128 x K16 = 64 tensor clocks each): the kernel is bound by the *issue rate* of tcgen05.mma, not by
// bandwidth.  The fix is to make every UMMA twice as large: the 128 output channels become the M dimension
// (weights = A operand) and the pixels the N dimension (activations = B operand), so one instruction covers
// N = 256 pixels (both 8x16 sub-tiles of the CTA's 8x32 region, which are contiguous rows of the same TMA box).
//
//   D^T[cout, pixel] = sum_k W[cout, k] * A[pixel, k]        M = 128 (cout), N = 128*SUBS (pixels), K = 16
//
// Consequences for the epilogue: a TMEM lane now holds ONE output channel for all pixels of the tile, so
//   * bias + time-embedding bias is one register per thread,
//   * the GroupNorm partial (sum, sum^2) of a channel is accumulated in registers -- no smem pass, no barriers,
//   * the fp16 NHWC tile is assembled in the swizzled staging buffer with 2-byte stores (64 pixels x 128 channels
//     per group, ping-pong) and leaves through TMA as before.
// Everything else (operand rings, dy/sub-tile reuse of the activation box, residual as identity K segment) is
// conv_tc2's.  SMEM operand reads per MMA clock drop from 128 B to 96 B as a side effect.
#include "kernels.h"

namespace sgmse {

CUtensorMap make_act_map(const void* p, int N, int H, int W, int C, int bw, int bh, int bn);
CUtensorMap make_w_map(const void* p, int Cout, int Ktot, int block_n);
int num_sms();

namespace {

constexpr int BLOCK_C = 128;                       // output channels per tile (UMMA M)
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 192;
constexpr int NUM_EPI_THREADS = 128;
constexpr int ROW_BYTES = 8 * 128;                 // one pixel row of the box (8 px x 64 ch fp16) = one swizzle atom
constexpr int W_BYTES = BLOCK_C * 128;             // weight tile: 128 cout x 64 cin fp16 = 16 KB
constexpr int GROUP_PX = 64;                       // pixels per epilogue group (8 pixel rows)
constexpr int GROUP_BYTES = 2 * GROUP_PX * 128;    // 64 px x 128 ch fp16 = 16 KB (two 64-channel chunks)
constexpr int MAX_SEG = 4;

struct Tc4Params {
  int tiles_w, tiles_h;
  int num_m_tiles, num_tiles, n_cblk;
  int N, Cout;
  int nseg;
  int seg_chunks[MAX_SEG];
  int seg_taps[MAX_SEG];
  int seg_kb0[MAX_SEG];
  const float* bias;
  const float* temb;
  int temb_stride;
  float scale;
  float* stats;
  int slots;
  int* dbg;
};

template <int SUBS, int A_STAGES, int B_STAGES>
struct Smem4 {
  static constexpr int A_ROWS = 16 * SUBS + 2;
  static constexpr int A_BYTES = A_ROWS * ROW_BYTES;
  static constexpr int OFF_W = A_STAGES * A_BYTES;
  static constexpr int OFF_STAGING = OFF_W + B_STAGES * W_BYTES;
  static constexpr int OFF_BARS = OFF_STAGING + 2 * GROUP_BYTES;
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
template <int SUBS>
__device__ __forceinline__ constexpr uint32_t idesc4() {      // f16 x f16 -> f32, M = 128, N = 128*SUBS
  return (1u << 4) | ((uint32_t)((128 * SUBS) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// one elected lane of a fully converged warp issues; all lanes carry identical (warp-uniform) operands
__device__ __forceinline__ void mma_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}
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
conv_tc4_kernel(const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_a1,
                const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_a3,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_d,
                const Tc4Params P) {
  using L = Smem4<SUBS, A_STAGES, B_STAGES>;
  constexpr int TILE_PX = 128 * SUBS;              // UMMA N
  constexpr int GROUPS = TILE_PX / GROUP_PX;
  constexpr uint32_t TMEM_COLS = 2 * TILE_PX;      // two accumulator stages
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BARS);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + A_STAGES;
  uint64_t* w_full = a_empty + A_STAGES;
  uint64_t* w_empty = w_full + B_STAGES;
  uint64_t* tmem_full = w_empty + B_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + L::OFF_TMEM_PTR);
  uint8_t* staging = smem + L::OFF_STAGING;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a0); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_d);
    if (P.nseg > 1) tma_prefetch_desc(&map_a1);
    if (P.nseg > 2) tma_prefetch_desc(&map_a2);
    if (P.nseg > 3) tma_prefetch_desc(&map_a3);
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();                    // everything above is on-chip set-up; global memory is first touched below
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_per_utt = P.tiles_w * P.tiles_h;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int c_blk = tile % P.n_cblk, m_tile = tile / P.n_cblk;   // channel tiles of one pixel tile run back to back (L2 reuse)
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
                mbar_wait(&w_empty[sb], pb ^ 1, P.dbg, 150 + sb);
                mbar_arrive_expect_tx(&w_full[sb], W_BYTES);
                tma_load_2d(smem + L::OFF_W + sb * W_BYTES, &map_w, &w_full[sb], kb * BLOCK_K, c_blk * BLOCK_C);
                if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
              }
            }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer: whole warp, warp-uniform control flow ===========================
    constexpr uint32_t IDESC = idesc4<SUBS>();
    int sa = 0; uint32_t pa = 0;
    int sb = 0; uint32_t pb = 0;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[as], as_phase ^ 1, P.dbg, 200 + as);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_PX);
      uint32_t acc = 0;
      for (int s = 0; s < P.nseg; ++s) {
        const int nd = P.seg_taps[s] == 9 ? 3 : 1;
        const int chunks = P.seg_chunks[s];
        for (int ch = 0; ch < chunks; ++ch)
          for (int dxi = 0; dxi < nd; ++dxi) {
            mbar_wait(&a_full[sa], pa, P.dbg, 300 + sa);
            const uint32_t a_base = smem_u32(smem + sa * L::A_BYTES);
            for (int dyi = 0; dyi < nd; ++dyi) {
              mbar_wait(&w_full[sb], pb, P.dbg, 350 + sb);
              tc_fence_after();
              const int row0 = nd == 3 ? dyi : 1;              // the box starts one pixel row above the tile
              const uint64_t wdesc = smem_desc_sw128(smem_u32(smem + L::OFF_W + sb * W_BYTES));       // A operand: weights
              const uint64_t pdesc = smem_desc_sw128(a_base + (uint32_t)(row0 * ROW_BYTES));          // B operand: pixels
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                mma_elect(d_tmem, wdesc + (uint64_t)(2 * k), pdesc + (uint64_t)(2 * k), IDESC, acc);
                acc = 1;
              }
              commit_elect(&w_empty[sb]);
              if (dyi == nd - 1) commit_elect(&a_empty[sa]);
              if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
            }
            if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          }
      }
      commit_elect(&tmem_full[as]);
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else {
    // =========================== epilogue (warps 2..5): thread = output channel ===========================
    const int e = threadIdx.x - 64;
    const int lg = warp & 3;
    const int ch = lg * 32 + lane;                 // TMEM lane = channel inside the tile
    // byte offset of (pixel row p, channel ch) inside a staged group, minus the swizzle term
    const int ch_chunk_off = (ch >> 6) * (GROUP_PX * 128) + (ch & 7) * 2;
    const int ch_c16 = (ch & 63) >> 3;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int c_blk = tile % P.n_cblk, m_tile = tile / P.n_cblk;   // channel tiles of one pixel tile run back to back (L2 reuse)
      const int n = m_tile / tiles_per_utt, rem = m_tile % tiles_per_utt;
      const int tx = rem % P.tiles_w, ty = rem / P.tiles_w;
      const int x0 = tx * 8, y0 = ty * (16 * SUBS);
      const int c_tile = c_blk * BLOCK_C;
      float bt = P.bias ? __ldg(P.bias + c_tile + ch) : 0.f;
      if (P.temb) bt += __ldg(P.temb + (size_t)n * P.temb_stride + c_tile + ch);
      float ssum = 0.f, ssq = 0.f;

      mbar_wait(&tmem_full[as], as_phase, P.dbg, 500 + as);
      tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < GROUPS; ++g) {
        uint8_t* buf = staging + (g & 1) * GROUP_BYTES;
        if (e == 0) tma_store_wait_read1();        // the store that last used `buf` has finished reading it
        named_bar_sync(1, NUM_EPI_THREADS);
        uint32_t r[64];
        tmem_ld_32x64(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(as * TILE_PX + g * GROUP_PX), r);
        tmem_ld_wait();
        if (g == GROUPS - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[as]);
        }
        uint8_t* base = buf + ch_chunk_off;
#pragma unroll
        for (int p = 0; p < GROUP_PX; ++p) {
          const __half h = __float2half_rn((__uint_as_float(r[p]) + bt) * P.scale);
          const float f = __half2float(h);
          ssum += f; ssq = fmaf(f, f, ssq);
          *reinterpret_cast<__half*>(base + p * 128 + ((ch_c16 ^ (p & 7)) << 4)) = h;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, NUM_EPI_THREADS);
        if (e == 0) {
          tma_store_4d(&map_d, buf, c_tile, x0, y0 + g * 8, n);
          tma_store_4d(&map_d, buf + GROUP_PX * 128, c_tile + 64, x0, y0 + g * 8, n);
          tma_store_commit();
        }
      }
      if (P.stats) {
        float2* o = reinterpret_cast<float2*>(P.stats + (((size_t)n * P.slots + (ty * P.tiles_w + tx)) * P.Cout + c_tile + ch) * 2);
        *o = make_float2(ssum, ssq);
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
    if (e == 0) tma_store_wait_all0();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int SUBS, int A_STAGES, int B_STAGES>
void launch4(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  using L = Smem4<SUBS, A_STAGES, B_STAGES>;
  Tc4Params P{};
  P.tiles_w = out.W / 8; P.tiles_h = out.H / (16 * SUBS);
  P.num_m_tiles = P.tiles_w * P.tiles_h * out.N;
  P.n_cblk = out.C / BLOCK_C;
  P.num_tiles = P.num_m_tiles * P.n_cblk;
  P.N = out.N; P.Cout = out.C;
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
  SG_CHECK(kb * 64 <= ld, "conv_tc4: K blocks (%d) exceed the packed weight row (%d)", kb * 64, ld);
  const CUtensorMap mw = make_w_map(a.w_tc, out.C, ld, BLOCK_C);
  const CUtensorMap md = make_act_map(out.p, out.N, out.H, out.W, out.C, 8, 8, 1);
  P.bias = a.bias; P.temb = a.temb; P.temb_stride = a.temb_stride;
  P.scale = a.scale;
  out.slots = P.tiles_w * P.tiles_h;
  P.stats = out.stats; P.slots = out.slots;
  P.dbg = dbg;
  auto kern = conv_tc4_kernel<SUBS, A_STAGES, B_STAGES>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs)) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
  }
  const int grid = P.num_tiles < num_sms() ? P.num_tiles : num_sms();
  launch_k(kern, dim3(grid), dim3(NUM_THREADS), (size_t)L::DYN_BYTES, st, ma[0], ma[1], ma[2], ma[3], mw, md, P);
  CUDA_OK(cudaGetLastError());
}

}  // namespace

bool conv_tc4_supported(const ConvArgs& a, const TensorDesc& out) {
  if (out.dt != DT_F16 || out.C % 128 != 0 || a.w_tc == nullptr) return false;
  if (out.W % 8 != 0 || out.H % 16 != 0) return false;
  for (int i = 0; i < a.nseg; ++i)
    if (a.seg[i].src.C % 64 != 0 || a.seg[i].src.dt != DT_F16) return false;
  if (a.residual && (!a.tc_identity_tail || a.nseg >= MAX_SEG || a.residual->C != out.C)) return false;
  return true;
}

void launch_conv_tc4(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  SG_CHECK(conv_tc4_supported(a, out), "conv_tc4: unsupported shape");
  // Ring depths trade prefetch distance against shared memory left for co-resident blocks of the HBM-bound kernels
  // that the other lane of the sampler graph runs concurrently (each needs 1 KB of reserved smem per block).
  if (out.H % 32 == 0) {
    launch4<2, 3, 5>(st, a, out, dbg);
  } else {
    launch4<1, 4, 6>(st, a, out, dbg);
  }
}

}  // namespace sgmse
