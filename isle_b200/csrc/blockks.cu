// blockks.cu -- restarted block Krylov-Schur for the top-k eigenpairs of B B^T, entirely on
// the device.  Restates the reference solver (block-ks/restarted_block_ks.h, SURVEY Appendix
// B.2) with the same numerical structure -- 3 block Gram-Schmidt passes per step, fp64
// rank-revealing MGS-QR with one re-orthogonalisation (block-ks/ks_utils.h:43-127), eig_sym
// of the projected matrix reading only its upper triangle, relative-residual stopping rule --
// but with the Krylov basis V (n x ncv, column-major) resident in HBM:
//   operator        spmm.cu (hand-written gather passes)
//   panel GEMMs     cuBLAS sgemm, fp32 FMA (plain library GEMMs: V^T F, F -= V H, V S)
//   MGS-QR          one cooperative kernel, fp64, 3 grid barriers per column
//   eig_sym         cuSOLVER ssyevd (upper), once per restart
// Only b+1 rows of H and one rank word travel to the host per restart / block step.
#include <cooperative_groups.h>

#include <cmath>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace isle {

static constexpr int kMaxB = 16;

// ------------------------------------------------------------------------------- helpers
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// arma::randu analogue (restarted_block_ks.h:212): uniform [0,1) floats, counter based.
__global__ void randu_kernel(float *__restrict__ out, size_t n, uint64_t seed, uint64_t stream_id)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t h = splitmix64(seed ^ splitmix64(stream_id * 0x100000001B3ull + i));
        out[i] = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
}

__global__ void add_block_kernel(float *__restrict__ dst, int ldd, const float *__restrict__ src, int lds,
                                 int rows, int cols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i % rows, cidx = i / rows;
    dst[r + (size_t)cidx * ldd] += src[r + (size_t)cidx * lds];
}

__global__ void zero_block_kernel(float *__restrict__ dst, int ldd, int rows, int cols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    dst[(i % rows) + (size_t)(i / rows) * ldd] = 0.0f;
}

// S[:, j] = T[:, mm-1-j], theta[j] = w[mm-1-j]   (ascending ssyevd output -> descending)
__global__ void reverse_top_kernel(const float *__restrict__ T, const float *__restrict__ w, int mm, int kk,
                                   float *__restrict__ S, float *__restrict__ theta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < mm * kk) {
        const int r = i % mm, j = i / mm;
        S[r + (size_t)j * mm] = T[r + (size_t)(mm - 1 - j) * mm];
    }
    if (i < kk) theta[i] = w[mm - 1 - i];
}

__global__ void set_diag_kernel(float *__restrict__ H, int ldh, int start, int count, const float *__restrict__ theta)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) H[(size_t)(start + i) * ldh + start + i] = theta[i];
}

// pack[j] = H[j,j] (j<k); pack[k + i + j*b] = H[k+i, j]
__global__ void pack_residual_kernel(const float *__restrict__ H, int ldh, int k, int b, float *__restrict__ pack)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < k) pack[i] = H[(size_t)i * ldh + i];
    if (i < k * b) {
        const int r = i % b, j = i / b;
        pack[k + i] = H[(size_t)j * ldh + k + r];
    }
}

// --------------------------------------------------------------------------- MGS-QR (fp64)
struct QrParams {
    const float *F;      // n x b, column-major, ld = n
    int64_t n;
    int b;
    double *a;           // n x b workspace
    double *q;           // n workspace
    float *Q;            // output columns (ld = n); only `rank` columns are written
    float *R;            // b x b column-major, pre-zeroed; row r = coefficients of pivot r
    int *rank_out;
    double *part;        // 2 x gridDim.x x kMaxB partial sums
};

__device__ __forceinline__ void block_partial(double (&acc)[kMaxB], int nb, double *__restrict__ dst,
                                              double (*s_red)[kMaxB])
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
    for (int j = 0; j < kMaxB; ++j) {
        if (j < nb) {
            double v = acc[j];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp][j] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x < nb) {
        double v = 0.0;
        for (int w = 0; w < nwarp; ++w) v += s_red[w][threadIdx.x];
        dst[(size_t)blockIdx.x * kMaxB + threadIdx.x] = v;
    }
    __syncthreads();
}

// every block sums the per-block partials in the same fixed order -> identical totals
__device__ __forceinline__ void grid_totals(const double *__restrict__ src, int nb, double *s_tot)
{
    if (threadIdx.x < nb) {
        double v = 0.0;
        for (unsigned blk = 0; blk < gridDim.x; ++blk) v += src[(size_t)blk * kMaxB + threadIdx.x];
        s_tot[threadIdx.x] = v;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256)
mgs_qr64_kernel(QrParams p)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_red[8][kMaxB];
    __shared__ double s_tot[kMaxB];
    __shared__ double s_bb[kMaxB];
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n = p.n;
    const int b = p.b;
    double *part0 = p.part, *part1 = p.part + (size_t)gridDim.x * kMaxB;
    double acc[kMaxB];

    // phase 0: widen to fp64 (ks_utils.h:58), partial ||a_0||^2
#pragma unroll
    for (int j = 0; j < kMaxB; ++j) acc[j] = 0.0;
    for (int64_t r = tid; r < n; r += nth) {
        for (int j = 0; j < b; ++j) {
            const double v = (double)p.F[r + (size_t)j * n];
            p.a[r + (size_t)j * n] = v;
            if (j == 0) acc[0] += v * v;
        }
    }
    block_partial(acc, 1, part0, s_red);
    int rank = 0;
    int flip = 0;   // buffer holding the current column-norm partials
    for (int i = 0; i < b; ++i) {
        grid.sync();
        grid_totals(flip ? part1 : part0, 1, s_tot);
        const float v_norm = (float)sqrt(s_tot[0]);          // ARMA_FPTYPE v_norm (ks_utils.h:66)
        double *nxt = flip ? part0 : part1;
        __syncthreads();
        if ((double)v_norm < 1e-6) {                          // discard column (ks_utils.h:69-72)
            acc[0] = 0.0;
            if (i + 1 < b)
                for (int64_t r = tid; r < n; r += nth) {
                    const double v = p.a[r + (size_t)(i + 1) * n];
                    acc[0] += v * v;
                }
            block_partial(acc, 1, nxt, s_red);
            flip ^= 1;
            continue;
        }
        const int nb = b - i;
        const double vn = (double)v_norm;
        // q = a_i / v_norm ; bb_j = q . a_{i+j}      (ks_utils.h:74-78)
#pragma unroll
        for (int j = 0; j < kMaxB; ++j) acc[j] = 0.0;
        for (int64_t r = tid; r < n; r += nth) {
            const double q = p.a[r + (size_t)i * n] / vn;
            p.q[r] = q;
            p.Q[r + (size_t)rank * n] = (float)q;
            for (int j = 0; j < nb; ++j) acc[j] += q * p.a[r + (size_t)(i + j) * n];
        }
        block_partial(acc, nb, nxt, s_red);
        grid.sync();
        grid_totals(nxt, nb, s_bb);
        // a_{i+j} -= q bb_j ; cc_j = q . a_{i+j}       (ks_utils.h:79-80)
        double *nxt2 = flip ? part1 : part0;
#pragma unroll
        for (int j = 0; j < kMaxB; ++j) acc[j] = 0.0;
        for (int64_t r = tid; r < n; r += nth) {
            const double q = p.q[r];
            for (int j = 0; j < nb; ++j) {
                const double v = p.a[r + (size_t)(i + j) * n] - q * s_bb[j];
                p.a[r + (size_t)(i + j) * n] = v;
                acc[j] += q * v;
            }
        }
        block_partial(acc, nb, nxt2, s_red);
        grid.sync();
        grid_totals(nxt2, nb, s_tot);
        // a_{i+j} -= q cc_j ; R(rank, i+j) = bb_j + cc_j ; partial ||a_{i+1}||^2   (ks_utils.h:81-82)
        if (blockIdx.x == 0 && threadIdx.x < nb)
            p.R[rank + (size_t)(i + threadIdx.x) * b] = (float)(s_bb[threadIdx.x] + s_tot[threadIdx.x]);
        acc[0] = 0.0;
        for (int64_t r = tid; r < n; r += nth) {
            const double q = p.q[r];
            for (int j = 0; j < nb; ++j) {
                const double v = p.a[r + (size_t)(i + j) * n] - q * s_tot[j];
                p.a[r + (size_t)(i + j) * n] = v;
                if (j == 1) acc[0] += v * v;
            }
        }
        // next column-norm partials go to the buffer the bb-totals were read from two
        // barriers ago (all readers are past it)
        block_partial(acc, 1, nxt, s_red);
        flip ^= 1;
        ++rank;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = rank;
}

struct KsState {
    Ctx &c;
    int64_t n;
    int k, b, ncv, m;
    DevBuf<float> V, H, F, C, Rb, Tm, Wev, S, theta, Vtmp, Htmp, pack, work;
    DevBuf<double> qa, qq, qpart;
    DevBuf<int> drank, dinfo;
    int lwork = 0;
    int qr_grid = 1;
    int H_rows = 0, H_cols = 0;
    uint64_t seed, rng_calls = 0;

    KsState(Ctx &ctx, int k_, int b_, uint64_t seed_) : c(ctx), n((int64_t)ctx.V), k(k_), b(b_), seed(seed_)
    {
        ncv = 2 * k + b;
        m = ncv - b;
        V.alloc((size_t)n * ncv);
        H.alloc((size_t)ncv * ncv);
        F.alloc((size_t)n * b);
        C.alloc((size_t)ncv * b);
        Rb.alloc((size_t)b * b);
        Tm.alloc((size_t)m * m);
        Wev.alloc(m);
        S.alloc((size_t)m * k);
        theta.alloc(k);
        Vtmp.alloc((size_t)n * k);
        Htmp.alloc((size_t)ncv * k);
        pack.alloc((size_t)(b + 1) * k);
        qa.alloc((size_t)n * b);
        qq.alloc((size_t)n);
        drank.alloc(1);
        dinfo.alloc(1);
        ISLE_CUDA_CHECK(cudaMemsetAsync(V.p, 0, V.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(H.p, 0, H.bytes(), c.stream));
        ISLE_CUSOLVER_CHECK(cusolverDnSsyevd_bufferSize(c.cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER,
                                                        m, Tm.p, m, Wev.p, &lwork));
        work.alloc((size_t)lwork);
        int per_sm = 0;
        ISLE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mgs_qr64_kernel, 256, 0));
        ISLE_REQUIRE(per_sm >= 1, ISLE_ERR_CUDA, "mgs_qr64_kernel cannot be made resident");
        int want = (int)((n + 255) / 256);
        qr_grid = std::max(1, std::min(want, c.num_sms));   // one CTA per SM at most
        qpart.alloc((size_t)2 * qr_grid * kMaxB);
    }

    float *Vcol(int j) { return V.p + (size_t)j * n; }
    float *Hat(int r, int col) { return H.p + (size_t)col * ncv + r; }

    void randu(float *dst, int cols)
    {
        randu_kernel<<<grid_for((size_t)n * cols, 256), 256, 0, c.stream>>>(dst, (size_t)n * cols, seed, ++rng_calls);
        count_launch(c);
    }

    // [Q,R] = mgs_qr64(Fsrc); Q -> Qdst (ld n); R -> Rb (b x b, zero padded).  Returns rank.
    int qr(const float *Fsrc, int cols, float *Qdst)
    {
        ISLE_CUDA_CHECK(cudaMemsetAsync(Rb.p, 0, Rb.bytes(), c.stream));
        QrParams p;
        p.F = Fsrc; p.n = n; p.b = cols; p.a = qa.p; p.q = qq.p; p.Q = Qdst; p.R = Rb.p;
        p.rank_out = drank.p; p.part = qpart.p;
        void *args[] = {&p};
        StatScope s(c, "ks_qr");
        ISLE_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mgs_qr64_kernel, dim3(qr_grid), dim3(256), args, 0, c.stream));
        count_launch(c);
        int rank = 0;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&rank, drank.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return rank;
    }
    // NOTE: for cols < b the R rows use leading dimension `cols`; callers that need R pass cols == b.

    void gemm_tn(int rows, int cols, const float *A, const float *Bm, float *Cm, int ldc)
    {   // Cm(rows x cols) = A(n x rows)^T Bm(n x cols)
        const float one = 1.f, zero = 0.f;
        ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_T, CUBLAS_OP_N, rows, cols, (int)n, &one, A, (int)n, Bm,
                                      (int)n, &zero, Cm, ldc));
        count_launch(c);
    }
    void gemm_sub(int rows, int cols, const float *A, const float *Hm, int ldh, float *Fm)
    {   // Fm(n x cols) -= A(n x rows) Hm(rows x cols)
        const float mone = -1.f, one = 1.f;
        ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, cols, rows, &mone, A, (int)n, Hm,
                                      ldh, &one, Fm, (int)n));
        count_launch(c);
    }

    // Orthogonalised random refill of V[:, nvecs:target)  (restarted_block_ks.h:106-131, 238-258)
    void refill(int nvecs, int target)
    {
        int tries = 0;
        while (nvecs < target && tries < 100) {
            ++tries;
            const int w = std::min(b, target - nvecs);
            randu(F.p, w);
            for (int pass = 0; pass < 2; ++pass) {
                gemm_tn(nvecs, w, V.p, F.p, C.p, ncv);
                gemm_sub(nvecs, w, V.p, C.p, ncv, F.p);
            }
            const int rk2 = qr(F.p, w, Vcol(nvecs));
            nvecs += rk2;
        }
        if (nvecs < target)   // leave the rest zero, as the reference's zero-padded V does
            ISLE_CUDA_CHECK(cudaMemsetAsync(Vcol(nvecs), 0, (size_t)(target - nvecs) * n * 4, c.stream));
    }

    void op(const float *X, float *Z)
    {
        StatScope s(c, "ks_op");
        spsptr_multiply_dev(c, b, X, Z);
    }

    // restarted_block_ks.h:204-259
    void init()
    {
        int rank = 0;
        do {
            randu(F.p, b);
            rank = qr(F.p, b, Vcol(0));
        } while (rank < b);
        op(Vcol(0), F.p);
        gemm_tn(b, b, V.p, F.p, Hat(0, 0), ncv);           // H = V0^T F
        gemm_sub(b, b, V.p, Hat(0, 0), ncv, F.p);          // F -= V0 H
        gemm_tn(b, b, V.p, F.p, C.p, ncv);                 // C = V0^T F
        add_block_kernel<<<(b * b + 255) / 256, 256, 0, c.stream>>>(Hat(0, 0), ncv, C.p, ncv, b, b);
        count_launch(c);
        gemm_sub(b, b, V.p, C.p, ncv, F.p);                // F -= V0 C
        rank = qr(F.p, b, Vcol(b));
        ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(b, 0), (size_t)ncv * 4, Rb.p, (size_t)b * 4, (size_t)b * 4, b,
                                          cudaMemcpyDeviceToDevice, c.stream));
        H_rows = 2 * b;
        H_cols = b;
        if (rank < b) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(Vcol(b + rank), 0, (size_t)(b - rank) * n * 4, c.stream));
            refill(b + rank, 2 * b);
        }
    }

    // restarted_block_ks.h:63-136
    void expand()
    {
        while (H_rows < ncv) {
            const int rows = H_rows, cols = H_cols;
            op(Vcol(cols), F.p);                                          // F = A V_k
            {
                StatScope s(c, "ks_orth", 6.0 * (double)n * rows * 4.0, 12.0 * (double)n * rows * b);
                float *Hk = Hat(0, cols);
                gemm_tn(rows, b, V.p, F.p, Hk, ncv);                      // Hk = W^T F
                gemm_sub(rows, b, V.p, Hk, ncv, F.p);                     // F -= W Hk
                for (int pass = 0; pass < 2; ++pass) {
                    gemm_tn(rows, b, V.p, F.p, C.p, ncv);                 // Ck = W^T F
                    gemm_sub(rows, b, V.p, C.p, ncv, F.p);                // F -= W Ck
                    add_block_kernel<<<(rows * b + 255) / 256, 256, 0, c.stream>>>(Hk, ncv, C.p, ncv, rows, b);
                    count_launch(c);
                }
                // new b rows of H are zero left of the R block
                zero_block_kernel<<<(b * (cols + b) + 255) / 256, 256, 0, c.stream>>>(Hat(rows, 0), ncv, b, cols + b);
                count_launch(c);
            }
            const int rk = qr(F.p, b, Vcol(rows));
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(rows, cols), (size_t)ncv * 4, Rb.p, (size_t)b * 4, (size_t)b * 4, b,
                                              cudaMemcpyDeviceToDevice, c.stream));
            H_rows += b;
            H_cols += b;
            if (rk < b) {
                ISLE_CUDA_CHECK(cudaMemsetAsync(Vcol(rows + rk), 0, (size_t)(b - rk) * n * 4, c.stream));
                refill(rows + rk, rows + b);
            }
        }
    }

    // restarted_block_ks.h:139-187
    void truncate(int nconv)
    {
        StatScope s(c, "ks_truncate");
        const int mm = m - nconv, kk = k - nconv;
        const float one = 1.f, zero = 0.f;
        ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Tm.p, (size_t)mm * 4, Hat(nconv, nconv), (size_t)ncv * 4, (size_t)mm * 4, mm,
                                          cudaMemcpyDeviceToDevice, c.stream));
        ISLE_CUSOLVER_CHECK(cusolverDnSsyevd(c.cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER, mm, Tm.p, mm,
                                             Wev.p, work.p, lwork, dinfo.p));
        count_launch(c);
        int info = 0;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&info, dinfo.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        ISLE_REQUIRE(info == 0, ISLE_ERR_CUDA, "evd(H) failed");   // restarted_block_ks.h:156-157
        reverse_top_kernel<<<(mm * kk + 255) / 256, 256, 0, c.stream>>>(Tm.p, Wev.p, mm, kk, S.p, theta.p);
        count_launch(c);
        // V_mid <- V[:, nconv:m) S ; new starts V[:, m:m+b) move to [k, k+b)
        ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, (int)n, kk, mm, &one, Vcol(nconv), (int)n,
                                      S.p, mm, &zero, Vtmp.p, (int)n));
        count_launch(c);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(Vcol(nconv), Vtmp.p, (size_t)n * kk * 4, cudaMemcpyDeviceToDevice, c.stream));
        ISLE_CUDA_CHECK(cudaMemcpyAsync(Vcol(k), Vcol(m), (size_t)n * b * 4, cudaMemcpyDeviceToDevice, c.stream));
        // residual coupling block: H[m:m+b, m-b:m) S[mm-b:mm, :]  -> Htmp rows [0,b)
        ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, b, kk, b, &one, Hat(m, m - b), ncv,
                                      S.p + (mm - b), mm, &zero, Htmp.p, ncv));
        count_launch(c);
        if (nconv > 0) {   // locked part: H[0:nconv, nconv:m) S -> Htmp rows [b, b+nconv)
            ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_N, CUBLAS_OP_N, nconv, kk, mm, &one, Hat(0, nconv), ncv,
                                          S.p, mm, &zero, Htmp.p + b, ncv));
            count_launch(c);
        }
        // rebuild H: columns >= nconv are cleared, then the three blocks are written back
        ISLE_CUDA_CHECK(cudaMemsetAsync(Hat(0, nconv), 0, (size_t)(ncv - nconv) * ncv * 4, c.stream));
        set_diag_kernel<<<(kk + 255) / 256, 256, 0, c.stream>>>(H.p, ncv, nconv, kk, theta.p);
        count_launch(c);
        ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(k, nconv), (size_t)ncv * 4, Htmp.p, (size_t)ncv * 4, (size_t)b * 4, kk,
                                          cudaMemcpyDeviceToDevice, c.stream));
        if (nconv > 0)
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(0, nconv), (size_t)ncv * 4, Htmp.p + b, (size_t)ncv * 4,
                                              (size_t)nconv * 4, kk, cudaMemcpyDeviceToDevice, c.stream));
        H_rows = k + b;
        H_cols = k;
    }

    // residual norms of the k Ritz pairs (restarted_block_ks.h:276-282); returns host copies
    void residuals(std::vector<float> &evs, std::vector<float> &norms)
    {
        pack_residual_kernel<<<(k * b + 255) / 256, 256, 0, c.stream>>>(H.p, ncv, k, b, pack.p);
        count_launch(c);
        std::vector<float> h((size_t)(b + 1) * k);
        ISLE_CUDA_CHECK(cudaMemcpyAsync(h.data(), pack.p, h.size() * 4, cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        evs.assign(h.begin(), h.begin() + k);
        norms.resize(k);
        for (int j = 0; j < k; ++j) {
            float s = 0.f;
            for (int i = 0; i < b; ++i) { const float v = h[(size_t)k + (size_t)j * b + i]; s += v * v; }
            norms[j] = std::sqrt(s);
        }
    }
};

void block_ks(Ctx &c, uint64_t k64, int b, int max_restarts, float tol, uint64_t seed, float *evalues_out,
              float *U_out, int *nconv_out)
{
    ISLE_REQUIRE(c.have_B, ISLE_ERR_ARG, "block_ks: build_B first");
    const int k = (int)k64;
    // restarted_block_ks.h:198: block size collapses to 1 when nev <= block size
    if (!(b < k)) b = 1;
    ISLE_REQUIRE(b >= 1 && b <= kMaxB, ISLE_ERR_ARG, "block_ks: block size must be in [1,16]");
    ISLE_REQUIRE(k >= 1 && (2 * k - b) % b == 0 && k % b == 0, ISLE_ERR_ARG,
                 "block_ks: k must be a multiple of the block size (reference expand() writes V out of "
                 "bounds otherwise, restarted_block_ks.h:71-100)");
    ISLE_REQUIRE((uint64_t)(2 * k + b) <= c.V, ISLE_ERR_ARG, "block_ks: ncv = 2k+b exceeds the vocabulary size");
    build_csr(c);
    KsState ks(c, k, b, seed);
    ks.init();
    int nconv = 0, n_restarts = 0;
    std::vector<float> evs, norms;
    ks.expand();
    while (n_restarts < max_restarts) {
        ks.truncate(nconv);
        ks.residuals(evs, norms);
        int first_bad = -1;
        for (int j = 0; j < k; ++j)
            if (norms[j] / evs[j] >= tol) { first_bad = j; break; }
        if (first_bad < 0) { nconv = k; break; }
        nconv = first_bad;
        ++n_restarts;
        ks.expand();
    }
    if (n_restarts == max_restarts) {   // restarted_block_ks.h:302-315 (norms not divided here)
        ks.residuals(evs, norms);
        int first_bad = -1;
        for (int j = 0; j < k; ++j)
            if (norms[j] >= tol) { first_bad = j; break; }
        nconv = first_bad < 0 ? k : first_bad;
    }
    nconv = std::min(nconv, k);
    c.counters["ks_restarts"] = n_restarts;
    c.counters["ks_nconv"] = nconv;

    // src/sparseMatrix.cpp:1209-1214: eigenvalues = diag(H)[0:k], eigenvectors = V[:, 0:k)
    if (evs.empty()) ks.residuals(evs, norms);
    if (evalues_out) std::copy(evs.begin(), evs.begin() + k, evalues_out);
    c.k = k64;
    c.U.alloc((size_t)c.V * k);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(c.U.p, ks.V.p, (size_t)c.V * k * 4, cudaMemcpyDeviceToDevice, c.stream));
    if (U_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(U_out, c.U.p, (size_t)c.V * k * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.have_U = true;
    c.have_P = false;
    if (nconv_out) *nconv_out = nconv;
    if (nconv != k) throw Error(ISLE_ERR_NOCONV, "block_ks: only " + std::to_string(nconv) + " of " + std::to_string(k) + " eigenpairs converged");
}

}  // namespace isle
