// tma_stream.cu -- how fast can ONE persistent block per SM stream fp64 boxes from HBM into shared memory with
// cp.async.bulk.tensor?  (No consumer work: the stage is released as soon as it has landed.)
// usage: tma_stream <bx> <by> <ops_per_stage> <stages> [x_start]     prints GB/s of reads
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar)), "r"(parity) : "memory");
}

__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ CUtensorMap map, int bx, int by, int ops, int stages, int x_start,
                                            int tiles_x, long long n_groups, double* sink, int pw, int nq)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int box_bytes = bx * by * 8;
    const int stage_bytes = ops * box_bytes;
    uint64_t* full = (uint64_t*) (smem + (size_t) stages * stage_bytes);
    uint64_t* empty = full + stages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(full + s)), "r"(pw < ops ? pw : ops));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(empty + s)));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int it = 0;
    if (warp < pw) {      // pw producer warps share the ops of a stage
        for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
            const int s = it % stages;
            wait(empty + s, ((it / stages) & 1) ^ 1);
            if (lane == 0) {
                int mine = 0;
                for (int o = warp; o < ops; o += pw) ++mine;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(full + s)), "r"(mine * box_bytes) : "memory");
                for (int o = warp; o < ops; o += pw) {
                    const long long t = nq > 1 ? g * (ops / nq) + o / nq : g * ops + o;
                    const int c0 = x_start + (int) (t % tiles_x) * bx, c1 = (int) (t / tiles_x) * by, c2 = o % nq;
                    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                                 ::"r"(s32(smem + (size_t) s * stage_bytes + o * box_bytes)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(s32(full + s)) : "memory");
                }
            }
            __syncwarp();
        }
    } else if (warp == pw) {
        double acc = 0.0;
        for (long long g = blockIdx.x; g < n_groups; g += gridDim.x, ++it) {
            const int s = it % stages;
            wait(full + s, (it / stages) & 1);
            acc += ((double*) (smem + (size_t) s * stage_bytes))[lane];
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(empty + s)) : "memory");
        }
        if (acc == 1.2345) sink[0] = acc;
    }
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const int bx = argc > 1 ? atoi(argv[1]) : 128, by = argc > 2 ? atoi(argv[2]) : 2, ops = argc > 3 ? atoi(argv[3]) : 19;
    const int stages = argc > 4 ? atoi(argv[4]) : 5, x_start = argc > 5 ? atoi(argv[5]) : 16;
    const int promo = argc > 6 ? atoi(argv[6]) : 0;
    const int pw = argc > 7 ? atoi(argv[7]) : 1;
    const int nq = argc > 8 ? atoi(argv[8]) : 1;
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr);
    Enc enc = (Enc) fp;
    const long long P = 528, rows = nq > 1 ? 514 * 514 : 1 << 20;       // 4.4 GB, or nq streams of 1.1 GB
    double *d, *sink;
    cudaMalloc(&d, P * rows * 8 * nq + 4096);
    cudaMalloc(&sink, 8);
    cudaMemset(d, 0, P * rows * 8 * nq);
    CUtensorMap map;
    cuuint64_t dims[3] = { (cuuint64_t) P, (cuuint64_t) rows, (cuuint64_t) nq }, strides[2] = { (cuuint64_t) P * 8, (cuuint64_t) P * rows * 8 };
    cuuint32_t box[3] = { (cuuint32_t) bx, (cuuint32_t) by, 1 }, es[3] = { 1, 1, 1 };
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : (promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int) r); return 1; }
    const int tiles_x = 512 / bx;
    const long long tiles = (long long) tiles_x * (rows / by), n_groups = nq > 1 ? tiles / (ops / nq) : tiles / ops;
    const int smem = stages * ops * bx * by * 8 + stages * 16;
    if (smem > 227 * 1024) { printf("bx=%d by=%d ops=%d stages=%d: %d bytes of shared memory do not fit\n", bx, by, ops, stages, smem); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        k<<<148, 32 * (pw + 1), smem>>>(map, bx, by, ops, stages, x_start, tiles_x, n_groups, sink, pw, nq);
        cudaEventRecord(b);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("FAILED %s\n", cudaGetErrorString(e)); return 1; }
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double bytes = (double) n_groups * ops * bx * by * 8;
    printf("box %3d x %2d (%5d B/op), %2d ops/stage, %d stages (%6d B in flight/SM), x_start %2d, promo %d, %d producer warps, %d streams: %7.1f GB/s  (%.3f ms)\n",
           bx, by, bx * by * 8, ops, stages, stages * ops * bx * by * 8, x_start, promo, pw, nq, bytes / best / 1e6, best);
    return 0;
}
