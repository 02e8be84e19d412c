// Multi-tensor Adam + weight re-packing in one pass over the parameters.
// torch.optim.Adam semantics (reference trainer.py:39-42): g' = g*scale + wd*p ; m = b1 m + (1-b1) g' ;
// v = b2 v + (1-b2) g'^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps).  The step counter lives on the
// device (advanced by a 1-thread kernel) so the whole update is CUDA-graph capturable; `scale` folds the
// 1/world_size of the data-parallel gradient all-reduce.  While the fp32 master value is in registers the
// bf16 hi(/lo) planes of both packed layouts (forward K-major, transposed) are rewritten, so the next step's
// tensor-core kernels need no separate repack pass.
#include "common.cuh"

namespace aclgan {

constexpr int kAdamChunk = 1024;

__global__ void adam_advance_kernel(float* hyper) {
    if (threadIdx.x == 0 && blockIdx.x == 0) hyper[6] += 1.0f;
}

__global__ void __launch_bounds__(256) adam_kernel(const aclgan_adam_tensor* __restrict__ table,
                                                   const int* __restrict__ chunks, const float* __restrict__ hyper) {
    const int tid = chunks[2 * blockIdx.x];
    const int64_t first = (int64_t)chunks[2 * blockIdx.x + 1] * kAdamChunk;
    const aclgan_adam_tensor T = table[tid];
    const int64_t numel = (int64_t)T.d[0] * T.d[1] * T.d[2] * T.d[3];
    const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], gscale = hyper[5];
    const double step = hyper[6];
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2s = (float)sqrt(1.0 - pow((double)b2, step));
    float* p = reinterpret_cast<float*>(T.p);
    float* m = reinterpret_cast<float*>(T.m);
    float* v = reinterpret_cast<float*>(T.v);
    const float* g = reinterpret_cast<const float*>(T.g) + T.goff;
#pragma unroll
    for (int i = 0; i < kAdamChunk / 256; ++i) {
        const int64_t e = first + i * 256 + threadIdx.x;
        if (e >= numel) break;
        int64_t r = e;
        const int kw = (int)(r % T.d[3]); r /= T.d[3];
        const int kh = (int)(r % T.d[2]); r /= T.d[2];
        const int ci = (int)(r % T.d[1]);
        const int co = (int)(r / T.d[1]);
        const float pv = p[e];
        const float gv = g[co * T.gs[0] + ci * T.gs[1] + kh * T.gs[2] + kw * T.gs[3]] * gscale + wd * pv;
        const float mv = b1 * m[e] + (1.f - b1) * gv;
        const float vv = b2 * v[e] + (1.f - b2) * gv * gv;
        const float np = pv - (lr / bc1) * mv / (sqrtf(vv) / bc2s + eps);
        m[e] = mv; v[e] = vv; p[e] = np;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            if (T.pk[k][0] == 0) continue;
            const int64_t o = T.aff[k][0] + co * T.aff[k][1] + ci * T.aff[k][2] + kh * T.aff[k][3] + kw * T.aff[k][4];
            const __nv_bfloat16 hi = __float2bfloat16_rn(np);
            reinterpret_cast<__nv_bfloat16*>(T.pk[k][0])[o] = hi;
            if (T.planes == 2)
                reinterpret_cast<__nv_bfloat16*>(T.pk[k][1])[o] = __float2bfloat16_rn(np - __bfloat162float(hi));
        }
    }
}

}  // namespace aclgan

extern "C" int aclgan_adam_advance(uint64_t hyper, void* stream) {
    aclgan::adam_advance_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<float*>(hyper));
    return (int)cudaGetLastError();
}

extern "C" int aclgan_adam_step(uint64_t table, uint64_t chunks, int32_t n_chunks, uint64_t hyper, void* stream) {
    if (n_chunks <= 0) return ACLGAN_OK;
    aclgan::adam_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const aclgan_adam_tensor*>(table), reinterpret_cast<const int*>(chunks),
        reinterpret_cast<const float*>(hyper));
    return (int)cudaGetLastError();
}
