// Probe of cp.async.bulk.tensor ... tile::gather4 on sm_100a: which tensor-map box shape it wants, how rows land in a
// SWIZZLE_128B destination, whether out-of-range row indices zero-fill and count towards the mbarrier tx bytes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/tma_gather_probe tools/tma_gather_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(const __grid_constant__ CUtensorMap tmap, int col, int r0, int r1, int r2, int r3, uint32_t tx_bytes,
                      float* out, int* status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    float* tile = reinterpret_cast<float*>(smem + (base - smem_u32(smem)));
    const uint32_t bar = base + 4096;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tile[i] = -7.0f;   // 4 KB sentinel
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(tx_bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(base), "l"(&tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
            : "memory");
        uint32_t done = 0;
        long spins = 0;
        while (!done && spins < 2000000) {
            asm volatile(
                "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(bar)
                : "memory");
            ++spins;
        }
        status[0] = (int)done;
        status[1] = (int)(spins > 0x7fffffff ? 0x7fffffff : spins);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    const int rows = 64, cols = 64;   // fp32 [rows, cols]; value = row * 1000 + col
    std::vector<float> h(rows * cols);
    for (int r = 0; r < rows; ++r)
        for (int c = 0; c < cols; ++c) h[r * cols + c] = r * 1000.f + c;
    float* d;
    CK(cudaMalloc(&d, h.size() * 4));
    CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    float* out;
    int* status;
    CK(cudaMalloc(&out, 4096));
    CK(cudaMalloc(&status, 16));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q));
    if (!encode) {
        printf("no cuTensorMapEncodeTiled\n");
        return 1;
    }
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
    struct Variant {
        const char* name;
        cuuint32_t box1;
        uint32_t tx;
        int r[4];
        int col;
    } variants[] = {
        {"box{32,1} tx=512 rows{3,1,10,63}", 1, 512, {3, 1, 10, 63}, 0},
        {"box{32,1} tx=512 rows{3,-1,10,64} (OOB zero fill?)", 1, 512, {3, -1, 10, 64}, 32},
        {"box{32,4} tx=512 rows{3,1,10,63}", 4, 512, {3, 1, 10, 63}, 0},
    };
    for (auto& v : variants) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
        cuuint64_t gstr[1] = {(cuuint64_t)cols * 4};
        cuuint32_t box[2] = {32, v.box1};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("=== %s : encode rc=%d\n", v.name, (int)rc);
        if (rc != CUDA_SUCCESS) continue;
        CK(cudaMemset(status, 0, 16));
        probe<<<1, 128, 8192>>>(tm, v.col, v.r[0], v.r[1], v.r[2], v.r[3], v.tx, out, status);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("kernel error: %s\n", cudaGetErrorString(e));
            return 2;   // sticky error: stop
        }
        int st[4];
        float ho[1024];
        CK(cudaMemcpy(st, status, 16, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ho, out, 4096, cudaMemcpyDeviceToHost));
        printf("barrier completed=%d spins=%d\n", st[0], st[1]);
        for (int row = 0; row < 8; ++row) {   // 8 smem rows of 128 B, print the first float of each 16 B chunk
            printf("smem row %d:", row);
            for (int c = 0; c < 8; ++c) printf(" %9.1f", ho[row * 32 + c * 4]);
            printf("\n");
        }
    }
    return 0;
}
