// tcgen05 implicit-GEMM 3x3 convolution with the GroupNorm-apply + SiLU of its input fused into the operand path.
//
// conv_tc4 reads a *normalised* activation tensor that a separate HBM-bound pass (gn_apply_plain_kernel: read x, write
// silu(a*x+b)) produced: 2 extra tensor passes per convolution, 22 % of a forward (profiles/r01_launches_*_v4).
// Here the 3x3 segment's pixel operand is produced by software instead of TMA: eight producer warps load the RAW
// tensor (x, or the two sources of a skip concat) with 128-bit loads, apply  y = silu(a[n,c]*x + b[n,c])  in registers
// and store the result straight into the SWIZZLE_128B operand stage the UMMA descriptors expect.  Shared-memory
// traffic is unchanged (the stores replace the TMA's writes); out-of-image pixels are written as zeros, i.e. the
// conv's zero padding applies after the activation, exactly like F.conv2d(silu(gn(x)), padding=1).
//   * silu(z) = hz*tanh(hz) + hz with hz = z/2, evaluated as tanh.approx.f16x2 (one MUFU op per two elements; the
//     activation is rounded to fp16 anyway) -- 8.7 k MUFU ops per 34 KB stage = 0.55 k clocks of the 1.5 k-clock stage.
//   * the TMA warp issues L2 prefetches (cp.async.bulk.prefetch.tensor) for the boxes two stages ahead so that the
//     producers' loads are L2 hits.
//   * raw 1x1 segments (Conv_2 shortcut inputs, identity residual) and the weights still arrive by TMA.
// Everything downstream of the operand rings is conv_tc4 (swapped operands, N = 256 pixel UMMAs, channel-per-lane
// epilogue).
#include "kernels.h"

namespace sgmse {

CUtensorMap make_act_map(const void* p, int N, int H, int W, int C, int bw, int bh, int bn);
CUtensorMap make_w_map(const void* p, int Cout, int Ktot, int block_n);
int num_sms();

namespace {

constexpr int BLOCK_C = 128;
constexpr int BLOCK_K = 64;
constexpr int UMMA_K = 16;
constexpr int NUM_PROD_GROUPS = 2;                 // producer groups alternate over the fused operand stages
constexpr int NUM_PROD_THREADS = 128;              // per group
constexpr int NUM_THREADS = 192 + NUM_PROD_GROUPS * NUM_PROD_THREADS;   // TMA, MMA, 4 epilogue warps, producers
constexpr int NUM_EPI_THREADS = 128;
constexpr int PREFETCH_DIST = 4;                   // L2 prefetch distance in operand stages
constexpr int SUBS = 2;
constexpr int ROW_BYTES = 8 * 128;
constexpr int A_ROWS = 16 * SUBS + 2;              // 34 pixel rows of 8 pixels
constexpr int A_BYTES = A_ROWS * ROW_BYTES;        // 34 KB
constexpr int A_PIXELS = A_ROWS * 8;               // 272
constexpr int ITEMS = A_PIXELS * 8 / NUM_PROD_THREADS;   // 17 16-byte vectors per producer thread per stage
constexpr int W_BYTES = BLOCK_C * 128;
constexpr int GROUP_PX = 64;
constexpr int GROUP_BYTES = 2 * GROUP_PX * 128;
constexpr int TILE_PX = 128 * SUBS;
constexpr int GROUPS = TILE_PX / GROUP_PX;

struct Tc5Params {
  int tiles_w, tiles_h;
  int num_m_tiles, num_tiles, n_cblk;
  int N, H, W, Cout;
  // segment 0: fused GroupNorm+SiLU 3x3 over concat(src0, src1)
  const __half* src0; const __half* src1;
  int C0, C1;
  const float2* ab;              // [N][C0+C1] (a, b)
  int chunks0;                   // (C0+C1)/64
  // segments 1..: raw 1x1 segments by TMA
  int nraw;
  int raw_chunks[2];
  int raw_kb0[2];
  const float* bias;
  const float* temb;
  int temb_stride;
  float scale;
  float* stats;
  int slots;
  int* dbg;
};

template <int A_STAGES, int B_STAGES>
struct Smem5 {
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
constexpr uint32_t IDESC5 = (1u << 4) | ((uint32_t)(TILE_PX >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // M128 N256

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
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// two activations: silu(z) = hz*tanh(hz) + hz with hz = z/2 given   (fp16 pair out)
__device__ __forceinline__ uint32_t silu_half_pair(float hz0, float hz1) {
  const __half2 hz = __floats2half2_rn(hz0, hz1);
  uint32_t hzu = *reinterpret_cast<const uint32_t*>(&hz), tu, yu;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(tu) : "r"(hzu));
  asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(yu) : "r"(hzu), "r"(tu));
  return yu;
}

template <int A_STAGES, int B_STAGES>
__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_tc5_kernel(const __grid_constant__ CUtensorMap map_x0, const __grid_constant__ CUtensorMap map_x1,
                const __grid_constant__ CUtensorMap map_r0, const __grid_constant__ CUtensorMap map_r1,
                const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_d,
                const Tc5Params P) {
  using L = Smem5<A_STAGES, B_STAGES>;
  constexpr uint32_t TMEM_COLS = 2 * TILE_PX;
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

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_x0); tma_prefetch_desc(&map_w); tma_prefetch_desc(&map_d);
    if (P.C1 > 0) tma_prefetch_desc(&map_x1);
    if (P.nraw > 0) tma_prefetch_desc(&map_r0);
    if (P.nraw > 1) tma_prefetch_desc(&map_r1);
    // every stage needs TWO arrivals on a_full: its filler's and the bystander role's (see conv_tc6.cu's header)
    for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 2); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < B_STAGES; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_per_utt = P.tiles_w * P.tiles_h;
  const int fused_stages = P.chunks0 * 3;            // A stages of segment 0 per tile: (chunk, dx)

  if (warp == 0) {
    // =========================== TMA producer: weights, raw 1x1 segments, L2 prefetch of the fused boxes ====
    if (lane == 0) {
      int sa = 0; uint32_t pa = 0;
      int sb = 0; uint32_t pb = 0;
      for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
        const int c_blk = tile % P.n_cblk, m_tile = tile / P.n_cblk;
        const int n = m_tile / tiles_per_utt, rem = m_tile % tiles_per_utt;
        const int x0 = (rem % P.tiles_w) * 8, y0 = (rem / P.tiles_w) * (16 * SUBS);
        auto prefetch_stage = [&](int i) {
          if (i >= fused_stages) return;
          const int ch = i / 3, dx = i % 3 - 1;
          const int cg = ch * 64;
          if (cg < P.C0) tma_prefetch_4d(&map_x0, cg, x0 + dx, y0 - 1, n);
          else tma_prefetch_4d(&map_x1, cg - P.C0, x0 + dx, y0 - 1, n);
        };
        for (int i = 0; i < PREFETCH_DIST; ++i) prefetch_stage(i);
        // ---- segment 0 (software-produced activations): only the weights come from here ----
        for (int i = 0; i < fused_stages; ++i) {
          prefetch_stage(i + PREFETCH_DIST);
          const int ch = i / 3, dxi = i % 3;
          // The stage itself belongs to the producer warps.  Still wait for its release: a role that skips phases of
          // a ring barrier can run two phases ahead, and a parity wait cannot tell phase k from phase k-2.
          mbar_wait(&a_empty[sa], pa ^ 1, P.dbg, 110 + sa);
          mbar_arrive(&a_full[sa]);                            // bystander arrival
          if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          for (int dyi = 0; dyi < 3; ++dyi) {
            const int kb = (dyi * 3 + dxi) * P.chunks0 + ch;
            mbar_wait(&w_empty[sb], pb ^ 1, P.dbg, 150 + sb);
            mbar_arrive_expect_tx(&w_full[sb], W_BYTES);
            tma_load_2d(smem + L::OFF_W + sb * W_BYTES, &map_w, &w_full[sb], kb * BLOCK_K, c_blk * BLOCK_C);
            if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
          }
        }
        // ---- raw 1x1 segments ----
        for (int s = 0; s < P.nraw; ++s) {
          const CUtensorMap* ma = s == 0 ? &map_r0 : &map_r1;
          for (int ch = 0; ch < P.raw_chunks[s]; ++ch) {
            mbar_wait(&a_empty[sa], pa ^ 1, P.dbg, 100 + sa);
            mbar_arrive_expect_tx(&a_full[sa], A_BYTES);
            tma_load_4d(smem + sa * A_BYTES, ma, &a_full[sa], ch * BLOCK_K, x0, y0 - 1, n);
            if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
            mbar_wait(&w_empty[sb], pb ^ 1, P.dbg, 160 + sb);
            mbar_arrive_expect_tx(&w_full[sb], W_BYTES);
            tma_load_2d(smem + L::OFF_W + sb * W_BYTES, &map_w, &w_full[sb], (P.raw_kb0[s] + ch) * BLOCK_K, c_blk * BLOCK_C);
            if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
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
      const uint32_t d_tmem = tmem_base + (uint32_t)(as * TILE_PX);
      uint32_t acc = 0;
      const int total_a = fused_stages + (P.nraw > 0 ? P.raw_chunks[0] : 0) + (P.nraw > 1 ? P.raw_chunks[1] : 0);
      for (int i = 0; i < total_a; ++i) {
        const int nd = i < fused_stages ? 3 : 1;
        mbar_wait(&a_full[sa], pa, P.dbg, 300 + sa);
        const uint32_t a_base = smem_u32(smem + sa * A_BYTES);
        for (int dyi = 0; dyi < nd; ++dyi) {
          mbar_wait(&w_full[sb], pb, P.dbg, 350 + sb);
          tc_fence_after();
          const int row0 = nd == 3 ? dyi : 1;
          const uint64_t wdesc = smem_desc_sw128(smem_u32(smem + L::OFF_W + sb * W_BYTES));
          const uint64_t pdesc = smem_desc_sw128(a_base + (uint32_t)(row0 * ROW_BYTES));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            mma_elect(d_tmem, wdesc + (uint64_t)(2 * k), pdesc + (uint64_t)(2 * k), IDESC5, acc);
            acc = 1;
          }
          commit_elect(&w_empty[sb]);
          if (dyi == nd - 1) commit_elect(&a_empty[sa]);
          if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
        }
        if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
      }
      commit_elect(&tmem_full[as]);
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else if (warp < 6) {
    // =========================== epilogue (warps 2..5): thread = output channel ===========================
    const int e = threadIdx.x - 64;
    const int lg = warp & 3;
    const int ch = lg * 32 + lane;
    const int ch_chunk_off = (ch >> 6) * (GROUP_PX * 128) + (ch & 7) * 2;
    const int ch_c16 = (ch & 63) >> 3;
    int as = 0; uint32_t as_phase = 0;
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int c_blk = tile % P.n_cblk, m_tile = tile / P.n_cblk;
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
        if (e == 0) tma_store_wait_read1();
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
  } else {
    // =========================== activation producers (warps 6..): raw x -> silu(a*x+b) -> operand stage ======
    // Two groups of four warps take alternate fused stages: while one group evaluates the activation of stage i the
    // other already has the loads of stage i+1 in flight.
    const int grp = (threadIdx.x - 192) / NUM_PROD_THREADS;
    const int pt = (threadIdx.x - 192) % NUM_PROD_THREADS;   // 0..127
    const int cv = pt & 7;                         // 8-channel vector inside the 64-channel chunk
    const int prow = pt >> 3;                      // 0..15
    const int Ct = P.C0 + P.C1;
    const int raw_stages = (P.nraw > 0 ? P.raw_chunks[0] : 0) + (P.nraw > 1 ? P.raw_chunks[1] : 0);
    int sa = 0; uint32_t pa = 0;
    uint32_t turn = 0;                             // running fused-stage counter (ownership = turn % groups)
    for (int tile = blockIdx.x; tile < P.num_tiles; tile += gridDim.x) {
      const int m_tile = tile / P.n_cblk;
      const int n = m_tile / tiles_per_utt, rem = m_tile % tiles_per_utt;
      const int x0 = (rem % P.tiles_w) * 8, y0 = (rem / P.tiles_w) * (16 * SUBS);
      for (int i = 0; i < fused_stages; ++i, ++turn) {
        if ((int)(turn % NUM_PROD_GROUPS) != grp) {          // the other group's stage
          if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
          continue;
        }
        const int ch = i / 3, dx = i % 3 - 1;
        const int cg = ch * 64 + cv * 8;           // channel of the concatenated input
        const __half* src; int Cs, cs;
        if (cg < P.C0) { src = P.src0; Cs = P.C0; cs = cg; } else { src = P.src1; Cs = P.C1; cs = cg - P.C0; }
        src += (size_t)n * P.H * P.W * Cs + cs;
        uint4 v[ITEMS];
        // all loads first (17 x 128 bit in flight per thread); out-of-image pixels stay zero
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          const int px = prow + 16 * j;
          const int y = y0 - 1 + (px >> 3), x = x0 + dx + (px & 7);
          v[j] = make_uint4(0u, 0u, 0u, 0u);
          if ((unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W)
            v[j] = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)y * P.W + x) * Cs));
        }
        float a[8], b[8];                          // (a, b)/2: the half argument of the tanh form of silu
        {
          const float4* q = reinterpret_cast<const float4*>(P.ab + (size_t)n * Ct + cg);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 w = __ldg(q + k);
            a[2 * k] = 0.5f * w.x; b[2 * k] = 0.5f * w.y; a[2 * k + 1] = 0.5f * w.z; b[2 * k + 1] = 0.5f * w.w;
          }
        }
        mbar_wait(&a_empty[sa], pa ^ 1, P.dbg, 600 + sa);
        uint8_t* stage = smem + sa * A_BYTES;
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
          const int px = prow + 16 * j;
          const int y = y0 - 1 + (px >> 3), x = x0 + dx + (px & 7);
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if ((unsigned)y < (unsigned)P.H && (unsigned)x < (unsigned)P.W) {
            const __half2* h = reinterpret_cast<const __half2*>(&v[j]);
            uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float2 f = __half22float2(h[k]);
              ow[k] = silu_half_pair(fmaf(a[2 * k], f.x, b[2 * k]), fmaf(a[2 * k + 1], f.y, b[2 * k + 1]));
            }
          }
          *reinterpret_cast<uint4*>(stage + px * 128 + ((cv ^ (px & 7)) << 4)) = o;
        }
        fence_proxy_async_smem();                  // generic-proxy stores -> visible to the tensor core's async proxy
        named_bar_sync(2 + grp, NUM_PROD_THREADS);
        if (pt == 0) mbar_arrive(&a_full[sa]);
        if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
      }
      // Stages of the raw segments belong to the TMA warp; wait for each release so that this role never runs two
      // phases ahead of a ring barrier (see the TMA warp).
      for (int i = 0; i < raw_stages; ++i) {
        mbar_wait(&a_empty[sa], pa ^ 1, P.dbg, 610 + sa);
        named_bar_sync(4, NUM_PROD_GROUPS * NUM_PROD_THREADS);
        if (grp == 0 && pt == 0) mbar_arrive(&a_full[sa]);    // bystander arrival: both groups have seen the release
        if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int A_STAGES, int B_STAGES>
void launch5(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  using L = Smem5<A_STAGES, B_STAGES>;
  Tc5Params P{};
  P.tiles_w = out.W / 8; P.tiles_h = out.H / (16 * SUBS);
  P.num_m_tiles = P.tiles_w * P.tiles_h * out.N;
  P.n_cblk = out.C / BLOCK_C;
  P.num_tiles = P.num_m_tiles * P.n_cblk;
  P.N = out.N; P.H = out.H; P.W = out.W; P.Cout = out.C;
  const TensorDesc& x0 = a.seg[0].src;
  const TensorDesc* x1 = a.gn_has_cat ? &a.gn_cat : nullptr;
  P.src0 = (const __half*)x0.p; P.C0 = x0.C;
  P.src1 = x1 ? (const __half*)x1->p : nullptr; P.C1 = x1 ? x1->C : 0;
  P.ab = a.gn_ab;
  P.chunks0 = (P.C0 + P.C1) / 64;
  // raw 1x1 segments: explicit ones first, then the residual (identity tail of the weights)
  const TensorDesc* raws[2] = {nullptr, nullptr};
  int nraw = 0;
  for (int i = 1; i < a.nseg; ++i) raws[nraw++] = &a.seg[i].src;
  if (a.residual) raws[nraw++] = a.residual;
  P.nraw = nraw;
  int kb = 9 * P.chunks0;
  for (int i = 0; i < nraw; ++i) { P.raw_chunks[i] = raws[i]->C / 64; P.raw_kb0[i] = kb; kb += P.raw_chunks[i]; }
  const int ld = a.w_tc_ld ? a.w_tc_ld : a.ktot();
  SG_CHECK(kb * 64 <= ld, "conv_tc5: K blocks (%d) exceed the packed weight row (%d)", kb * 64, ld);
  const CUtensorMap mx0 = make_act_map(x0.p, x0.N, x0.H, x0.W, x0.C, 8, A_ROWS, 1);
  const CUtensorMap mx1 = x1 ? make_act_map(x1->p, x1->N, x1->H, x1->W, x1->C, 8, A_ROWS, 1) : mx0;
  const CUtensorMap mr0 = raws[0] ? make_act_map(raws[0]->p, raws[0]->N, raws[0]->H, raws[0]->W, raws[0]->C, 8, A_ROWS, 1) : mx0;
  const CUtensorMap mr1 = raws[1] ? make_act_map(raws[1]->p, raws[1]->N, raws[1]->H, raws[1]->W, raws[1]->C, 8, A_ROWS, 1) : mx0;
  const CUtensorMap mw = make_w_map(a.w_tc, out.C, ld, BLOCK_C);
  const CUtensorMap md = make_act_map(out.p, out.N, out.H, out.W, out.C, 8, 8, 1);
  P.bias = a.bias; P.temb = a.temb; P.temb_stride = a.temb_stride;
  P.scale = a.scale;
  out.slots = P.tiles_w * P.tiles_h;
  P.stats = out.stats; P.slots = out.slots;
  P.dbg = dbg;
  auto kern = conv_tc5_kernel<A_STAGES, B_STAGES>;
  static unsigned long long attr_devs = 0;
  if (first_use_on_device(attr_devs)) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::DYN_BYTES));
  }
  const int grid = P.num_tiles < num_sms() ? P.num_tiles : num_sms();
  kern<<<grid, NUM_THREADS, L::DYN_BYTES, st>>>(mx0, mx1, mr0, mr1, mw, md, P);
  CUDA_OK(cudaGetLastError());
}

}  // namespace

// Shape test usable before the tensors exist: a 3x3 conv over (c0 [+ c1]) raw channels at H x W, cout outputs,
// `nraw` extra raw 1x1 segments (incl. the residual).
bool conv_tc5_shape_ok(int H, int W, int c0, int c1, int cout, int nraw) {
  return H % 32 == 0 && W % 8 == 0 && c0 % 64 == 0 && c1 % 64 == 0 && cout % 128 == 0 && nraw <= 2;
}

void launch_conv_tc5(cudaStream_t st, const ConvArgs& a, TensorDesc& out, int* dbg) {
  SG_CHECK(a.gn_ab != nullptr && a.nseg >= 1 && a.seg[0].taps == 9, "conv_tc5: needs a fused GroupNorm 3x3 segment");
  const int nraw = (a.nseg - 1) + (a.residual ? 1 : 0);
  SG_CHECK(conv_tc5_shape_ok(out.H, out.W, a.seg[0].src.C, a.gn_has_cat ? a.gn_cat.C : 0, out.C, nraw) &&
               out.dt == DT_F16 && a.w_tc != nullptr && (!a.residual || a.tc_identity_tail),
           "conv_tc5: unsupported shape");
  launch5<3, 5>(st, a, out, dbg);
}

}  // namespace sgmse
