#include "common.cuh"

namespace aclgan {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_tmap(const aclgan_tmap_spec* s, CUtensorMap* out) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return ACLGAN_ERR_DRIVER;
    if (s->rank < 2 || s->rank > 5 || s->elem_bytes != 2) return ACLGAN_ERR_SHAPE;
    if (s->base & 15) return ACLGAN_ERR_ALIGN;
    cuuint64_t dims[5], strides[4];
    cuuint32_t box[5], estr[5];
    for (uint32_t i = 0; i < s->rank; ++i) {
        dims[i] = s->dims[i];
        box[i] = s->box[i];
        estr[i] = 1;
        if (box[i] < 1 || box[i] > 256 || dims[i] < 1) return ACLGAN_ERR_SHAPE;
        if (i > 0) {
            strides[i - 1] = s->strides[i];
            if (s->strides[i] & 15) return ACLGAN_ERR_ALIGN;
        }
    }
    if (box[0] * s->elem_bytes != 128) return ACLGAN_ERR_SHAPE;   // one 128B swizzle row per box row
    CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, s->rank, reinterpret_cast<void*>(s->base), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "aclgan: cuTensorMapEncodeTiled failed (%d): rank %u dims %llu %llu %llu %llu box %u %u %u %u\n",
                (int)r, s->rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
                (unsigned long long)(s->rank > 2 ? dims[2] : 0), (unsigned long long)(s->rank > 3 ? dims[3] : 0), box[0],
                box[1], s->rank > 2 ? box[2] : 0, s->rank > 3 ? box[3] : 0);
        return 1000 + (int)r;
    }
    return ACLGAN_OK;
}

int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();      // no device (host-only plan building): assume a B200
            n = 148;
        }
    }
    return n;
}

}  // namespace aclgan

extern "C" int aclgan_version(void) { return ACLGAN_ABI_VERSION; }
extern "C" const char* aclgan_build_info(void) {
    return "aclgan_b200 sm_100a (tcgen05 + TMA implicit GEMM), built " __DATE__ " " __TIME__;
}

// Makes `stream` wait for an event recorded OUTSIDE a stream capture (cudaEventWaitExternal): inside a capture it becomes
// an external event-wait node of the graph, outside it is a plain cudaStreamWaitEvent.  Used by trainer.gen_update: its
// captured graph starts the generator passes while the previous dis_update (graph + Adam on another stream) is still
// running and only waits for it right before the discriminator passes.
extern "C" int aclgan_stream_wait_external_event(void* stream, void* event) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaError_t e = cudaStreamIsCapturing((cudaStream_t)stream, &st);
    if (e != cudaSuccess) return (int)e;
    // (the external flag is only valid while capturing)
    const unsigned flags = st == cudaStreamCaptureStatusActive ? cudaEventWaitExternal : cudaEventWaitDefault;
    return (int)cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, flags);
}
