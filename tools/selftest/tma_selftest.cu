// tma_selftest.cu -- which cp.async.bulk.tensor forms does this driver/GPU accept for fp64 lattices?
// usage: tma_selftest <variant>   (one variant per process: an illegal instruction kills the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, int box_elems, double* out, int issuer_warp, int lanes)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bar = (uint64_t*) (smem + 65536);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == issuer_warp) {
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(box_elems * 8 * lanes) : "memory");
        __syncwarp();
        if (lane < lanes) {
            double* dst = (double*) smem + lane * box_elems;
            if constexpr (RANK == 2)
                asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                             ::"r"(s32(dst)), "l"(&map), "r"(c0), "r"(c1 + lane), "r"(s32(bar)) : "memory");
            else
                asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                             ::"r"(s32(dst)), "l"(&map), "r"(c0), "r"(c1), "r"(c2), "r"(c3 + lane), "r"(s32(bar)) : "memory");
        }
    }
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar)) : "memory");
    for (int i = threadIdx.x; i < box_elems * lanes; i += blockDim.x) out[i] = ((double*) smem)[i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                        const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv)
{
    const char v = argc > 1 ? argv[1][0] : 'A';
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) { printf("%c: no encoder\n", v); return 1; }
    Enc enc = (Enc) fp;
    const int P = 48, Y = 14, Z = 12, Q = 15;
    const long long qstride = (long long) P * Y * Z + 16;
    const size_t n = (size_t) qstride * Q;
    std::vector<double> h(n);
    for (size_t i = 0; i < n; ++i) h[i] = (double) i;
    double *d, *out;
    cudaMalloc(&d, n * 8);
    cudaMalloc(&out, 65536);
    cudaMemcpy(d, h.data(), n * 8, cudaMemcpyHostToDevice);
    CUtensorMap map;
    int rank = (v == 'A' || v == 'B') ? 2 : 4;
    int c[4] = { 0, 0, 0, 0 };
    int lanes = 1, warp = 0, threads = 32;
    cuuint64_t dims[4] = { P, Y, Z, Q };
    cuuint64_t strides[3] = { (cuuint64_t) P * 8, (cuuint64_t) P * Y * 8, (cuuint64_t) qstride * 8 };
    cuuint32_t box[4] = { 32, 8, 1, 1 }, es[4] = { 1, 1, 1, 1 };
    if (v == 'B') c[0] = 1;
    if (v == 'C' || v == 'D' || v == 'E' || v == 'F') { c[0] = 15; c[1] = 1; c[2] = 1; c[3] = 3; }
    if (v == 'G') { c[0] = 16; c[1] = 1; c[2] = 1; c[3] = 3; }        // 4-D, even start
    if (v == 'H') { c[0] = 14; c[1] = 1; c[2] = 1; c[3] = 3; }        // 4-D, 16-byte aligned start
    if (v == 'D' || v == 'E') lanes = 8;
    if (v == 'E') { threads = 544; warp = 16; }
    if (v == 'F') { dims[3] = 1; dims[2] = (cuuint64_t) Z; rank = 3; }   // not used below
    if (v == 'Y') { // the d3q19 / 512^3 shape
        c[0] = 15; c[1] = 1; c[2] = 1; c[3] = 3;
    }
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, rank == 2 ? 2 : 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%c: encode failed %d\n", v, (int) r); return 1; }
    const int smem = 65536 + 64;
    if (rank == 2) {
        cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<2><<<1, threads, smem>>>(map, c[0], c[1], c[2], c[3], 256, out, warp, lanes);
    } else {
        cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        k<4><<<1, threads, smem>>>(map, c[0], c[1], c[2], c[3], 256, out, warp, lanes);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%c: FAILED %s\n", v, cudaGetErrorString(e)); return 1; }
    std::vector<double> o(256 * lanes);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int l = 0; l < lanes; ++l)
        for (int y = 0; y < 8; ++y)
            for (int x = 0; x < 32; ++x) {
                long long gx = c[0] + x, gy = c[1] + y + (rank == 2 ? l : 0), gz = c[2], gq = c[3] + (rank == 2 ? 0 : l);
                double want = (gx < P && gy < Y) ? (double) (gq * qstride + gz * (long long) P * Y + gy * P + gx) : 0.0;
                if (o[l * 256 + y * 32 + x] != want) ++bad;
            }
    printf("%c: ok, %d wrong values\n", v, bad);
    return 0;
}
