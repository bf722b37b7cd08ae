// Single-head global self-attention over the H*W tokens of one utterance (AttnBlockpp core).
//
// Reference: /root/reference/sgmse/backbones/ncsnpp_utils/layerspp.py:82-86
//   w = softmax_k( sum_c q[c,tok] k[c,key] / sqrt(C) );  h[tok] = sum_key w v[key]
// 0.16 % of the network FLOPs -> fp32 CUDA-core kernel, smem-tiled; the q/k/v
// projections and the output projection run through the convolution kernels as 1x1 convs.
#include "kernels.h"

namespace sgmse {

namespace {
constexpr int KT = 64;  // keys per tile

template <typename T, int QT>
__global__ void __launch_bounds__(128) attention_kernel(const T* __restrict__ qkv, int S, int C, float scale,
                                                        T* __restrict__ out) {
  extern __shared__ float sm[];
  const int ldq = C + 1;
  const int lds = S + 1;
  float* Qs = sm;                    // [QT][C+1]
  float* KVs = Qs + QT * ldq;        // [KT][C+1]
  float* Ss = KVs + KT * ldq;        // [QT][S+1]
  const int n = blockIdx.y;
  const int q0 = blockIdx.x * QT;
  const int tid = threadIdx.x;
  const T* base = qkv + (size_t)n * S * 3 * C;
  constexpr int QG = QT / 4;         // query groups of 4
  constexpr int NT = 128 / QG;       // threads along keys / channels
  const int tq = tid / NT, tk = tid % NT;

  // ---- load Q tile ----
  for (int i = tid; i < QT * (C / 8); i += 128) {
    const int q = i / (C / 8), cv = i % (C / 8);
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (q0 + q < S) { Vec8<T> v; v.load(base + (size_t)(q0 + q) * 3 * C + cv * 8); v.get(f); }
#pragma unroll
    for (int j = 0; j < 8; ++j) Qs[q * ldq + cv * 8 + j] = f[j];
  }
  // ---- scores ----
  for (int k0 = 0; k0 < S; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * (C / 8); i += 128) {
      const int k = i / (C / 8), cv = i % (C / 8);
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (k0 + k < S) { Vec8<T> v; v.load(base + (size_t)(k0 + k) * 3 * C + C + cv * 8); v.get(f); }
#pragma unroll
      for (int j = 0; j < 8; ++j) KVs[k * ldq + cv * 8 + j] = f[j];
    }
    __syncthreads();
    // each thread: 4 queries x (KT/NT) keys, keys interleaved by NT
    constexpr int KPT = KT / NT;
    float acc[4][KPT];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < KPT; ++j) acc[i][j] = 0.f;
    for (int c = 0; c < C; ++c) {
      float qv[4], kv[KPT];
#pragma unroll
      for (int i = 0; i < 4; ++i) qv[i] = Qs[(tq * 4 + i) * ldq + c];
#pragma unroll
      for (int j = 0; j < KPT; ++j) kv[j] = KVs[(tk + j * NT) * ldq + c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < KPT; ++j) acc[i][j] = fmaf(qv[i], kv[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const int k = k0 + tk + j * NT;
        if (k < S) Ss[(tq * 4 + i) * lds + k] = acc[i][j] * scale;
      }
  }
  __syncthreads();
  // ---- softmax rows (one warp per row, round robin) ----
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < QT; q += 4) {
      float* row = Ss + q * lds;
      float mx = -INFINITY;
      for (int k = lane; k < S; k += 32) mx = fmaxf(mx, row[k]);
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int k = lane; k < S; k += 32) { const float e = __expf(row[k] - mx); row[k] = e; sum += e; }
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float inv = 1.f / sum;
      for (int k = lane; k < S; k += 32) row[k] *= inv;
    }
  }
  // ---- O = P V ; thread: 4 queries x channels {tk, tk+NT, ...} handled in passes of 16 channels ----
  for (int cb = 0; cb < C; cb += 16 * NT) {
    float o[4][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j) o[i][j] = 0.f;
    for (int k0 = 0; k0 < S; k0 += KT) {
      __syncthreads();
      for (int i = tid; i < KT * (C / 8); i += 128) {
        const int k = i / (C / 8), cv = i % (C / 8);
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (k0 + k < S) { Vec8<T> v; v.load(base + (size_t)(k0 + k) * 3 * C + 2 * C + cv * 8); v.get(f); }
#pragma unroll
        for (int j = 0; j < 8; ++j) KVs[k * ldq + cv * 8 + j] = f[j];
      }
      __syncthreads();
      const int kmax = (S - k0) < KT ? (S - k0) : KT;
      for (int k = 0; k < kmax; ++k) {
        float p[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = Ss[(tq * 4 + i) * lds + k0 + k];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int c = cb + tk + j * NT;
          const float v = c < C ? KVs[k * ldq + c] : 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i][j] = fmaf(p[i], v, o[i][j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = q0 + tq * 4 + i;
      if (q >= S) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int c = cb + tk + j * NT;
        if (c < C) Act<T>::st(out + ((size_t)n * S + q) * C + c, o[i][j]);
      }
    }
  }
}

template <typename T, int QT>
static void run(cudaStream_t st, const TensorDesc& qkv, TensorDesc& out, int S, int C, size_t smem) {
  auto kern = attention_kernel<T, QT>;
  CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(cdiv(S, QT), qkv.N);
  kern<<<grid, 128, smem, st>>>((const T*)qkv.p, S, C, 1.0f / sqrtf((float)C), (T*)out.p);
  CUDA_OK(cudaGetLastError());
}
}  // namespace

void launch_attention(cudaStream_t st, const TensorDesc& qkv, TensorDesc& out) {
  const int C = out.C, S = qkv.H * qkv.W;
  SG_CHECK(qkv.C == 3 * C && C % 8 == 0, "attention: qkv must have 3C channels");
  auto smem_for = [&](int QT) { return (size_t)((QT + KT) * (C + 1) + QT * (S + 1)) * sizeof(float); };
  const size_t lim = 220 * 1024;
  int QT = 32;
  if (smem_for(32) > lim) QT = 16;
  if (QT == 16 && smem_for(16) > lim) QT = 8;
  SG_CHECK(smem_for(QT) <= lim, "attention: %d tokens x %d channels does not fit the v1 kernel", S, C);
#define GO(T) \
  do { if (QT == 32) run<T, 32>(st, qkv, out, S, C, smem_for(32)); \
       else if (QT == 16) run<T, 16>(st, qkv, out, S, C, smem_for(16)); \
       else run<T, 8>(st, qkv, out, S, C, smem_for(8)); } while (0)
  if (qkv.dt == DT_F16) GO(__half); else GO(float);
#undef GO
}

}  // namespace sgmse
