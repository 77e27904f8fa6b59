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
static constexpr int kCholDynSmem = 16 * 256 * 8;   // cholqr2_kernel reduction scratch

// ------------------------------------------------------------------------------- helpers
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

// arma::randu analogue (restarted_block_ks.h:212): uniform [0,1) floats, counter based.  Fills rows [r0, r0 + nl) of
// a global ng x cols column-major matrix into a local nl x cols one (ld nl): every rank of a row-sharded basis draws
// its slice of the same matrix a single GPU would draw.
__global__ void randu_kernel(float *__restrict__ out, size_t nl, size_t ng, size_t r0, int cols, uint64_t seed, uint64_t stream_id)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nl * (size_t)cols; i += stride) {
        const size_t r = i % nl, cc = i / nl;
        const uint64_t h = splitmix64(seed ^ splitmix64(stream_id * 0x100000001B3ull + (r0 + r) + cc * ng));
        out[i] = (float)(h >> 40) * (1.0f / 16777216.0f);
    }
}

// loc(nl x cols, ld nl) = rows [r0, r0 + nl) of full(ng x cols, ld ng)
__global__ void slice_rows_kernel(const float *__restrict__ full, size_t ng, size_t r0, size_t nl, int cols, float *__restrict__ loc)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < nl * (size_t)cols; i += stride) loc[i] = full[r0 + i % nl + (i / nl) * ng];
}

// full(ng x cols, ld ng) <- the ranks' row slices, gathered as [rank][cols][base] (rank g holds rows [g base, ...))
__global__ void place_rows_kernel(const float *__restrict__ gathered, size_t ng, size_t base, int world, int cols, float *__restrict__ full)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < ng * (size_t)cols; i += stride) {
        const size_t r = i % ng, cc = i / ng, g = r / base;
        full[i] = gathered[(g * cols + cc) * base + (r - g * base)];
    }
}

// pad(base x cols) <- loc(nl x cols, ld nl), zero rows beyond nl
__global__ void pad_rows_kernel(const float *__restrict__ loc, size_t nl, size_t base, int cols, float *__restrict__ pad)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < base * (size_t)cols; i += stride) {
        const size_t r = i % base, cc = i / base;
        pad[i] = r < nl ? loc[r + cc * nl] : 0.0f;
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


// ------------------------------------------------------------------- CholeskyQR2 (fp64)
// Fast path of compute_qr (block-ks/ks_utils.h:43-127) for a numerically full-rank block:
//   G1 = F^T F, R1 = chol(G1);  G2 = (F R1^-1)^T (F R1^-1), R2 = chol(G2);
//   Q = F (R1^-1 R2^-1) rounded to fp32, R = R2 R1,
// all in fp64 with fixed-order grid reductions (bitwise reproducible, so replicated ranks
// stay in lock step).  One cooperative launch, two grid barriers.  When a Cholesky pivot
// falls under the reference's rank cut (column norm < 1e-6, ks_utils.h:69) or under the
// fp64 noise floor of the Gram matrix, rank_out = -1 and the caller runs the rank-revealing
// MGS kernel above, which follows the reference step by step.
struct CholQrParams {
    const float *F;      // n x b, column-major, ld = n
    int64_t n;
    int b;
    float *Q;            // n x b out (ld = n)
    float *R;            // b x b column-major out
    int *rank_out;       // b, or -1 = fall back to mgs_qr64_kernel
    double *part;        // gridDim.x x 256 partial Gram sums
};

// Gram partials of this block's rows of X = F * T (T upper triangular b x b in smem, or identity when
// T == nullptr).  The CTA walks 64-row tiles: loads are coalesced (thread = row (tid & 63) of the tile, columns
// tid >> 6 + {0, 4, 8, 12}: a warp reads 32 consecutive rows of one column) and the next tile is prefetched
// into registers while this one is reduced; in the reduction thread (p = tid & 15, r = tid >> 4) owns rows
// r, r + 16, r + 32, r + 48 of the tile and column p of the Gram matrix.
static constexpr int kCholTile = 64;

__device__ __forceinline__ void gram_partial(const CholQrParams &p, const double (*T)[kMaxB], double (*xs)[kMaxB + 1],
                                             double (*ys)[kMaxB + 1], double *red, double *dst)
{
    const int pc = threadIdx.x & 15, r = threadIdx.x >> 4;
    const int lrow = threadIdx.x & (kCholTile - 1), lcg = threadIdx.x >> 6;
    const int b = p.b;
    double acc[kMaxB];
#pragma unroll
    for (int q = 0; q < kMaxB; ++q) acc[q] = 0.0;
    const int64_t step = (int64_t)gridDim.x * kCholTile;
    int64_t row0 = (int64_t)blockIdx.x * kCholTile;
    float nxt[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int col = lcg + 4 * u;
        nxt[u] = (row0 + lrow < p.n && col < b) ? p.F[row0 + lrow + (size_t)col * p.n] : 0.f;
    }
    for (; row0 < p.n; row0 += step) {
#pragma unroll
        for (int u = 0; u < 4; ++u) xs[lrow][lcg + 4 * u] = (double)nxt[u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // prefetch the next tile while this one is reduced
            const int col = lcg + 4 * u;
            const int64_t rown = row0 + step + lrow;
            nxt[u] = (rown < p.n && col < b) ? p.F[rown + (size_t)col * p.n] : 0.f;
        }
        __syncthreads();
        double (*src)[kMaxB + 1] = xs;
        if (T) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int rr = r + 16 * u;
                double y = 0.0;
                for (int i = 0; i <= pc; ++i) y += xs[rr][i] * T[i][pc];
                ys[rr][pc] = (pc < b) ? y : 0.0;
            }
            __syncthreads();
            src = ys;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int rr = r + 16 * u;
            const double mine = src[rr][pc];
#pragma unroll
            for (int q = 0; q < kMaxB; ++q) acc[q] += mine * src[rr][q];
        }
        __syncthreads();
    }
    // reduce over the 16 row-owners r in a fixed order
#pragma unroll
    for (int q = 0; q < kMaxB; ++q) red[(size_t)r * 256 + pc * 16 + q] = acc[q];
    __syncthreads();
    double v = 0.0;
    for (int rr = 0; rr < 16; ++rr) v += red[(size_t)rr * 256 + threadIdx.x];
    dst[(size_t)blockIdx.x * 256 + threadIdx.x] = v;
    __syncthreads();
}

// fixed-order sum of nb per-block partial Gram matrices (thread t < 256 owns entry t); the loads are independent of the
// running sum, so unrolling keeps eight in flight
__device__ __forceinline__ double sum_partials(const double *part, unsigned nb)
{
    double v = 0.0;
    unsigned blk = 0;
    for (; blk + 8 <= nb; blk += 8) {
        double t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t[u] = part[(size_t)(blk + u) * 256 + threadIdx.x];
#pragma unroll
        for (int u = 0; u < 8; ++u) v += t[u];
    }
    for (; blk < nb; ++blk) v += part[(size_t)blk * 256 + threadIdx.x];
    return v;
}

// Every block: G = sum of partials (fixed order) or the given Gram matrix, Cholesky R (upper), Rinv = R^-1.  Returns
// false when a pivot is not safely positive.
__device__ __forceinline__ bool chol_and_invert(const double *part, int b, double *G, double (*R)[kMaxB],
                                                double (*Rinv)[kMaxB], int *s_ok, bool apply_rank_cut, bool part_is_G = false)
{
    {
        G[threadIdx.x] = part_is_G ? part[threadIdx.x] : sum_partials(part, gridDim.x);
        R[threadIdx.x >> 4][threadIdx.x & 15] = 0.0;
        Rinv[threadIdx.x >> 4][threadIdx.x & 15] = 0.0;
        if (threadIdx.x == 0) *s_ok = 1;
    }
    __syncthreads();
    // right-looking Cholesky on the first warp: lane c owns column c of the upper factor
    if (threadIdx.x < 32) {
        const int c = threadIdx.x;
        for (int j = 0; j < b; ++j) {
            double d = G[j * 16 + j];
            for (int i = 0; i < j; ++i) d -= R[i][j] * R[i][j];
            // reference rank cut: ||a_j - proj||_2 < 1e-6 (ks_utils.h:69); Gram noise floor ~ 1e-9 G_jj
            const double floor_ = apply_rank_cut ? fmax(1e-12, 1e-9 * G[j * 16 + j]) : 0.0;
            if (!(d > floor_)) {            // uniform across the warp: every lane evaluates the same pivot
                if (c == 0) *s_ok = 0;
                break;
            }
            const double rjj = sqrt(d);
            if (c == j) R[j][j] = rjj;
            if (c > j && c < b) {
                double t = G[j * 16 + c];
                for (int i = 0; i < j; ++i) t -= R[i][j] * R[i][c];
                R[j][c] = t / rjj;
            }
            __syncwarp();
        }
        __syncwarp();
        if (*s_ok && c < b) {   // column c of the upper-triangular inverse by back substitution
            Rinv[c][c] = 1.0 / R[c][c];
            for (int i = c - 1; i >= 0; --i) {
                double t = 0.0;
                for (int l = i + 1; l <= c; ++l) t += R[i][l] * Rinv[l][c];
                Rinv[i][c] = -t / R[i][i];
            }
        }
    }
    __syncthreads();
    return *s_ok != 0;
}

__global__ void __launch_bounds__(256)
cholqr2_kernel(CholQrParams p)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ double xs[kCholTile][kMaxB + 1], ys[kCholTile][kMaxB + 1];
    __shared__ double G[256];
    __shared__ double R1[kMaxB][kMaxB], R1i[kMaxB][kMaxB], R2[kMaxB][kMaxB], R2i[kMaxB][kMaxB], T[kMaxB][kMaxB];
    __shared__ int s_ok;
    extern __shared__ double red[];   // 16 x 256 doubles
    const int b = p.b;
    double *part0 = p.part, *part1 = p.part + (size_t)gridDim.x * 256;

    gram_partial(p, nullptr, xs, ys, red, part0);
    grid.sync();
    if (!chol_and_invert(part0, b, G, R1, R1i, &s_ok, true)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = -1;
        return;   // uniform across the grid: every block computed the same decision
    }
    gram_partial(p, R1i, xs, ys, red, part1);
    grid.sync();
    if (!chol_and_invert(part1, b, G, R2, R2i, &s_ok, false)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = -1;
        return;
    }
    // T = R1^-1 R2^-1 ; R = R2 R1
    {
        const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
        double t = 0.0, rr = 0.0;
        for (int l = 0; l < kMaxB; ++l) { t += R1i[i][l] * R2i[l][j]; rr += R2[i][l] * R1[l][j]; }
        T[i][j] = t;
        if (blockIdx.x == 0 && i < b && j < b) p.R[i + (size_t)j * b] = (float)rr;
    }
    __syncthreads();
    // Q = F T: same 64-row tiles, coalesced loads and stores (thread = row of the tile, four columns each)
    const int lrow = threadIdx.x & (kCholTile - 1), lcg = threadIdx.x >> 6;
    for (int64_t row0 = (int64_t)blockIdx.x * kCholTile; row0 < p.n; row0 += (int64_t)gridDim.x * kCholTile) {
        const int64_t row = row0 + lrow;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = lcg + 4 * u;
            xs[lrow][col] = (row < p.n && col < b) ? (double)p.F[row + (size_t)col * p.n] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = lcg + 4 * u;
            if (row < p.n && col < b) {
                double y = 0.0;
                for (int i = 0; i <= col; ++i) y += xs[lrow][i] * T[i][col];
                p.Q[row + (size_t)col * p.n] = (float)y;
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = b;
}

// ---- the same for a row-sharded block (SURVEY 8e option B): every rank holds rows [r0, r1) of F, the two Gram
// matrices are all-reduced (256 doubles), every rank factorises them identically and writes its rows of Q.
//   gram_rows_kernel (T = identity) -> sum -> all-reduce G1 -> gram_rows_kernel (T = R1^-1) -> sum -> all-reduce G2 ->
//   cholqr2_finish_kernel.  A bad pivot in either factorisation makes finish report rank -1 (every rank alike).
__global__ void __launch_bounds__(256)
gram_rows_kernel(CholQrParams p, const double *__restrict__ G1 /* NULL: plain Gram of F */)
{
    __shared__ double xs[kCholTile][kMaxB + 1], ys[kCholTile][kMaxB + 1];
    __shared__ double G[256];
    __shared__ double R1[kMaxB][kMaxB], R1i[kMaxB][kMaxB];
    __shared__ int s_ok;
    extern __shared__ double red[];   // 16 x 256 doubles
    if (G1) {
        if (!chol_and_invert(G1, p.b, G, R1, R1i, &s_ok, true, true)) return;     // finish reports the failure
        gram_partial(p, R1i, xs, ys, red, p.part);
    } else {
        gram_partial(p, nullptr, xs, ys, red, p.part);
    }
}

__global__ void __launch_bounds__(256)
sum_partials_kernel(const double *__restrict__ part, unsigned nb, double *__restrict__ G)
{
    G[threadIdx.x] = sum_partials(part, nb);
}

__global__ void __launch_bounds__(256)
cholqr2_finish_kernel(CholQrParams p, const double *__restrict__ G1, const double *__restrict__ G2)
{
    __shared__ double xs[kCholTile][kMaxB + 1];
    __shared__ double G[256];
    __shared__ double R1[kMaxB][kMaxB], R1i[kMaxB][kMaxB], R2[kMaxB][kMaxB], R2i[kMaxB][kMaxB], T[kMaxB][kMaxB];
    __shared__ int s_ok;
    const int b = p.b;
    bool ok = chol_and_invert(G1, b, G, R1, R1i, &s_ok, true, true);
    __syncthreads();
    ok = ok && chol_and_invert(G2, b, G, R2, R2i, &s_ok, false, true);
    if (!ok) {
        if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = -1;
        return;
    }
    {
        const int i = threadIdx.x >> 4, j = threadIdx.x & 15;
        double t = 0.0, rr = 0.0;
        for (int l = 0; l < kMaxB; ++l) { t += R1i[i][l] * R2i[l][j]; rr += R2[i][l] * R1[l][j]; }
        T[i][j] = t;
        if (blockIdx.x == 0 && i < b && j < b) p.R[i + (size_t)j * b] = (float)rr;
    }
    __syncthreads();
    const int lrow = threadIdx.x & (kCholTile - 1), lcg = threadIdx.x >> 6;
    for (int64_t row0 = (int64_t)blockIdx.x * kCholTile; row0 < p.n; row0 += (int64_t)gridDim.x * kCholTile) {
        const int64_t row = row0 + lrow;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = lcg + 4 * u;
            xs[lrow][col] = (row < p.n && col < b) ? (double)p.F[row + (size_t)col * p.n] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int col = lcg + 4 * u;
            if (row < p.n && col < b) {
                double y = 0.0;
                for (int i = 0; i <= col; ++i) y += xs[lrow][i] * T[i][col];
                p.Q[row + (size_t)col * p.n] = (float)y;
            }
        }
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.rank_out = b;
}

// ------------------------------------------------------------- tall-skinny panel products
// The two block Gram-Schmidt primitives of expand() (restarted_block_ks.h:83-90) for a basis
// W (n x rows, column-major, ld n) and a block F (n x b): HBM/L2-bound streams over W with
// N = b <= 16, for which a library GEMM spends its time in split-K bookkeeping.
//   wtf:   C = W^T F   per-chunk partials (fixed order -> deterministic) + ordered reduce
//   fsub:  F -= W C,   one thread per row of F, C staged in shared memory
static constexpr int kWtfChunk = 512;    // rows of W per CTA (its F slice is staged in shared memory)
static constexpr int kWtfCols = 32;      // columns of W per CTA (4 per warp)

template <int BP>
__global__ void __launch_bounds__(256, 2)
wtf_partial_kernel(const float *__restrict__ W, const float *__restrict__ F, int64_t n, int rows, int b,
                   float *__restrict__ partial /* [chunks][rows][BP] */)
{
    // the CTA's slice of F sits in shared memory ([column][row]: lanes read consecutive words), so the inner
    // loop issues only the W loads: four columns x four 32-row slabs = 16 independent loads per lane in flight
    __shared__ float sF[BP][kWtfChunk];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.y * kWtfChunk;
    for (int t = threadIdx.x; t < BP * kWtfChunk; t += 256) {
        const int c = t / kWtfChunk, r = t % kWtfChunk;
        sF[c][r] = (c < b && i0 + r < n) ? F[i0 + r + (size_t)c * n] : 0.f;
    }
    __syncthreads();
    const int j0 = blockIdx.x * kWtfCols + warp * 4;
    if (j0 >= rows) return;
    float acc[4][BP];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
        for (int c = 0; c < BP; ++c) acc[jj][c] = 0.f;
    const int nj = min(4, rows - j0);
    const float *wp[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) wp[jj] = W + (size_t)(j0 + min(jj, nj - 1)) * n;   // clamp: unused columns re-read a valid one
#pragma unroll 1
    for (int it = 0; it < kWtfChunk / 32; it += 4) {
        float w[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + (it + u) * 32 + lane;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) w[u][jj] = i < n ? __ldg(wp[jj] + i) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = (it + u) * 32 + lane;
#pragma unroll
            for (int c = 0; c < BP; ++c) {
                const float f = sF[c][r];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[jj][c] = fmaf(w[u][jj], f, acc[jj][c]);
            }
        }
    }
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
#pragma unroll
        for (int c = 0; c < BP; ++c) {
            float v = acc[jj][c];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[jj][c] = v;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            if (jj < nj) {
#pragma unroll
                for (int c = 0; c < BP; ++c)
                    partial[((size_t)blockIdx.y * rows + (j0 + jj)) * BP + c] = acc[jj][c];
            }
    }
}

// C[j + c ldc] = sum over chunks in a fixed order (lane-strided partial sums, then a shuffle tree: the
// same association on every run and every rank); optionally Hk[j + c ldh] (+)= the same value.
// One warp per (j, c).
template <int BP>
__global__ void __launch_bounds__(256)
wtf_reduce_kernel(const float *__restrict__ partial, int chunks, int rows, int b,
                  float *__restrict__ C, int ldc, float *__restrict__ Hk, int ldh, int hk_assign)
{
    const int lane = threadIdx.x & 31;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (t >= rows * BP) return;
    const int j = t / BP, c = t % BP;
    float v = 0.f;
    for (int ch = lane; ch < chunks; ch += 32) v += partial[((size_t)ch * rows + j) * BP + c];
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0 && c < b) {
        C[j + (size_t)c * ldc] = v;
        if (Hk) Hk[j + (size_t)c * ldh] = hk_assign ? v : Hk[j + (size_t)c * ldh] + v;
    }
}

template <int BP>
__global__ void __launch_bounds__(256)
fsub_kernel(const float *__restrict__ W, const float *__restrict__ C, int ldc, int64_t n, int rows, int b,
            float *__restrict__ F)
{
    constexpr int JT = 256;
    __shared__ __align__(16) float cs[JT][BP];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float acc[BP];
#pragma unroll
    for (int c = 0; c < BP; ++c) acc[c] = 0.f;
    for (int jt = 0; jt < rows; jt += JT) {
        const int nj = min(JT, rows - jt);
        __syncthreads();
        for (int t = threadIdx.x; t < nj * BP; t += blockDim.x) {
            const int j = t / BP, c = t % BP;
            cs[j][c] = (c < b) ? C[(jt + j) + (size_t)c * ldc] : 0.f;
        }
        __syncthreads();
        if (i < n) {
            const float *wp = W + i + (size_t)jt * n;
            int j = 0;
            for (; j + 16 <= nj; j += 16) {
                float w[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) w[u] = __ldg(wp + (size_t)(j + u) * n);
#pragma unroll
                for (int u = 0; u < 16; ++u) {
#pragma unroll
                    for (int q = 0; q < BP / 4; ++q) {
                        const float4 cv = *reinterpret_cast<const float4 *>(&cs[j + u][4 * q]);
                        acc[4 * q] = fmaf(w[u], cv.x, acc[4 * q]);
                        acc[4 * q + 1] = fmaf(w[u], cv.y, acc[4 * q + 1]);
                        acc[4 * q + 2] = fmaf(w[u], cv.z, acc[4 * q + 2]);
                        acc[4 * q + 3] = fmaf(w[u], cv.w, acc[4 * q + 3]);
                    }
                }
            }
            for (; j < nj; ++j) {
                const float w = __ldg(wp + (size_t)j * n);
#pragma unroll
                for (int c = 0; c < BP; ++c) acc[c] = fmaf(w, cs[j][c], acc[c]);
            }
        }
    }
    if (i < n) {
#pragma unroll
        for (int c = 0; c < BP; ++c)
            if (c < b) F[i + (size_t)c * n] -= acc[c];
    }
}

// ---- vector-load variant of wtf for 16-byte aligned columns (n % 4 == 0): 128 KB in flight per SM.
// Same arithmetic and fixed-order reductions, different tiling.  Measured (profiles/r1h): 2.25 vs 2.15 TB/s at
// k = 2000 -- both variants are bound by the FMA / shared-memory rate of a 6 flop/byte SIMT product, not by
// the bytes in flight (an 8-group vector variant of fsub was slower than the scalar kernel and was dropped).
static constexpr int kWtf2Chunk = 1024;   // rows of W per CTA = 8 slabs of 128 rows (32 lanes x float4)
static constexpr int kWtf2Cols = 32;      // columns of W per CTA: 16 warps x 2

template <int BP>
__global__ void __launch_bounds__(512, 1)
wtf_partial_v2_kernel(const float *__restrict__ W, const float *__restrict__ F, int64_t n, int rows, int b,
                      float *__restrict__ partial /* [chunks][rows][BP] */)
{
    extern __shared__ float4 sF4[];       // [BP][256] float4 = this CTA's 1024-row slice of F, column by column
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i0 = (int64_t)blockIdx.y * kWtf2Chunk;
    const int j0 = blockIdx.x * kWtf2Cols + warp * 2;
    const int nj = min(2, rows - j0);     // <= 0: this warp only helps staging F
    // 1. all 16 W loads of this warp go out first (8 KB in flight per warp)
    float4 w[2][8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int64_t i = i0 + u * 128 + lane * 4;
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
            w[jj][u] = (i < n && jj < nj) ? __ldg(reinterpret_cast<const float4 *>(W + (size_t)(j0 + jj) * n + i))
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // 2. F slice -> shared memory
    for (int t = threadIdx.x; t < BP * 256; t += 512) {
        const int c = t >> 8, r4 = t & 255;
        const int64_t i = i0 + r4 * 4;
        sF4[t] = (c < b && i < n) ? __ldg(reinterpret_cast<const float4 *>(F + (size_t)c * n + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    if (nj <= 0) return;
    float acc[2][BP];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int c = 0; c < BP; ++c) acc[jj][c] = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int c = 0; c < BP; ++c) {
            const float4 f = sF4[c * 256 + u * 32 + lane];
#pragma unroll
            for (int jj = 0; jj < 2; ++jj)
                acc[jj][c] = fmaf(w[jj][u].x, f.x, fmaf(w[jj][u].y, f.y, fmaf(w[jj][u].z, f.z, fmaf(w[jj][u].w, f.w, acc[jj][c]))));
        }
    }
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
        for (int c = 0; c < BP; ++c) {
            float v = acc[jj][c];
#pragma unroll
            for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            acc[jj][c] = v;
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int jj = 0; jj < 2; ++jj)
            if (jj < nj) {
#pragma unroll
                for (int c = 0; c < BP; ++c)
                    partial[((size_t)blockIdx.y * rows + (j0 + jj)) * BP + c] = acc[jj][c];
            }
    }
}

// The fp32 FMA panel engine: W^T F and F -= W C for a basis W of n rows (ld n) and at most ncv columns.
struct PanelSimt {
    Ctx &c;
    int64_t n = 0;
    int ncv = 0;
    bool vec = true;                   // vector-load W^T F when the columns are 16-byte aligned
    DevBuf<float> wpart;
    explicit PanelSimt(Ctx &ctx) : c(ctx) {}
    void init(int64_t n_, int ncv_, bool vec_)
    {
        n = n_; ncv = ncv_; vec = vec_;
        const int chunks = (int)((n + kWtfChunk - 1) / kWtfChunk);
        wpart.alloc((size_t)chunks * ncv * kMaxB);
    }

    // Cm(rows x b, ld ldc) = W^T Fm with W = V[:, 0:rows); when Hk != nullptr, Hk (ld ncv) is assigned
    // (first pass) or incremented (correction passes) by the same coefficients.
    template <int BP>
    void wtf_t(const float *W, int rows, int cols, const float *Fm, float *Cm, int ldc, float *Hk, bool assign)
    {
        const bool v2 = vec && n % 4 == 0;
        const int chunk_rows = v2 ? kWtf2Chunk : kWtfChunk;
        const int chunks = (int)((n + chunk_rows - 1) / chunk_rows);
        {
            StatScope s(c, "ks_wtf", (double)n * rows * 4.0, 2.0 * (double)n * rows * cols);
            if (v2) {
                const int smem = BP * 256 * (int)sizeof(float4);
                // per launch, not once per process: the attribute is per device
                ISLE_CUDA_CHECK(cudaFuncSetAttribute(wtf_partial_v2_kernel<BP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                dim3 g((rows + kWtf2Cols - 1) / kWtf2Cols, chunks);
                wtf_partial_v2_kernel<BP><<<g, 512, smem, c.stream>>>(W, Fm, n, rows, cols, wpart.p);
            } else {
                dim3 g((rows + kWtfCols - 1) / kWtfCols, chunks);
                wtf_partial_kernel<BP><<<g, 256, 0, c.stream>>>(W, Fm, n, rows, cols, wpart.p);
            }
        }
        {
            StatScope s(c, "ks_wtfred", (double)chunks * rows * BP * 4.0);
            wtf_reduce_kernel<BP><<<(rows * BP * 32 + 255) / 256, 256, 0, c.stream>>>(wpart.p, chunks, rows, cols, Cm, ldc, Hk,
                                                                                       ncv, assign ? 1 : 0);
        }
        count_launch(c, 2);
    }
    template <int BP>
    void fsub_t(const float *W, int rows, int cols, const float *Cm, int ldc, float *Fm)
    {
        StatScope s(c, "ks_fsub", (double)n * rows * 4.0, 2.0 * (double)n * rows * cols);
        fsub_kernel<BP><<<(unsigned)((n + 255) / 256), 256, 0, c.stream>>>(W, Cm, ldc, n, rows, cols, Fm);
        count_launch(c);
    }
    void wtf(const float *W, int rows, int cols, const float *Fm, float *Cm, int ldc, float *Hk, bool assign)
    {
        if (cols <= 4) wtf_t<4>(W, rows, cols, Fm, Cm, ldc, Hk, assign);
        else if (cols <= 8) wtf_t<8>(W, rows, cols, Fm, Cm, ldc, Hk, assign);
        else if (cols <= 12) wtf_t<12>(W, rows, cols, Fm, Cm, ldc, Hk, assign);
        else wtf_t<16>(W, rows, cols, Fm, Cm, ldc, Hk, assign);
    }
    void fsub(const float *W, int rows, int cols, const float *Cm, int ldc, float *Fm)
    {
        if (cols <= 4) fsub_t<4>(W, rows, cols, Cm, ldc, Fm);
        else if (cols <= 8) fsub_t<8>(W, rows, cols, Cm, ldc, Fm);
        else if (cols <= 12) fsub_t<12>(W, rows, cols, Cm, ldc, Fm);
        else fsub_t<16>(W, rows, cols, Cm, ldc, Fm);
    }
};

struct KsState {
    Ctx &c;
    int64_t n;                 // rows of the basis THIS rank holds (= ng unless the basis is row-sharded)
    int64_t ng;                // rows of the operator (vocabulary size)
    // Row-sharded basis (SURVEY 8e option B, document-sharded contexts): rank g holds rows [g base, min(ng, (g+1) base))
    // of V and F.  Panel products, QR Gram sums and the truncation GEMM then cost 1/world per rank; what travels per block
    // step: the next block all-gathered for the operator (ng x b), the operator's all-reduce (as before), the rows x b
    // Gram-Schmidt coefficients all-reduced per pass, two 16 x 16 Gram matrices.  H and every decision stay replicated
    // and identical on all ranks (all-reduced inputs, deterministic kernels).
    bool shard = false;
    int64_t r0 = 0, base = 0;
    DevBuf<float> Xfull, Zfull, Qfull, gpad, grecv, ctmp;
    DevBuf<double> gram1, gram2;
    int k, b, ncv, m;
    DevBuf<float> V, H, F, C, Rb, Tm, Wev, S, theta, Vtmp, Htmp, pack, work;
    PanelSimt ps;
    DevBuf<double> qa, qq, qpart, cpart;
    int chol_grid = 1, gs_passes = 3;
    bool fast_qr = true, custom_orth = true, panel_v2 = true, panel_tc = true;
    int panel_tc_min_rows = 0;
    bool gs_elide = true;
    PanelTc ptc;
    DevBuf<int> drank, dinfo;
    bool defer_rank = true;
    int force_fallback = 0, fast_qr_calls = 0;
    DevBuf<float> F2;          // second operator-output buffer (deferred rank check)
    int *h_rank = nullptr, *h_rank_dev = nullptr;     // host-mapped: written by the fast QR kernels, read after ev_rank
    cudaEvent_t ev_rank = nullptr;
    int lwork = 0;
    int qr_grid = 1;
    int H_rows = 0, H_cols = 0;
    uint64_t seed, rng_calls = 0;

    ~KsState()
    {
        if (h_rank) cudaFreeHost(h_rank);
        if (ev_rank) cudaEventDestroy(ev_rank);
    }
    KsState(const KsState &) = delete;
    KsState(Ctx &ctx, int k_, int b_, uint64_t seed_) : c(ctx), n((int64_t)ctx.V), ng((int64_t)ctx.V), k(k_), b(b_), ps(ctx), seed(seed_)
    {
        ncv = 2 * k + b;
        m = ncv - b;
        // option ks_row_shard: 1 on, 0 off, -1 (default) on when it pays: the extra small collectives (a gather, three
        // coefficient all-reduces and two Gram all-reduces per block step) cost ~0.1 ms per step, which a basis that
        // streams from HBM (>= 256 MB: k >= ~300 at a 100k vocabulary) repays many times and an L2-resident one does not
        // (c2, k = 100, 2 GPUs: 53.4 ms per step sharded against 49 replicated)
        const int rs = c.opt("ks_row_shard", -1);
        const bool rs_on = rs > 0 || (rs < 0 && (double)ng * ncv * 4.0 >= 256e6);
        if (c.world > 1 && rs_on && c.opt("ks_panel_tc", 1) != 0 && c.opt("ks_custom_orth", 1) != 0 &&
            ng % 4 == 0 && b <= 16) {
            base = ((ng + c.world - 1) / c.world + 3) / 4 * 4;
            if ((int64_t)(c.world - 1) * base + 64 <= ng) {      // every rank gets a real slice
                shard = true;
                r0 = (int64_t)c.rank * base;
                n = std::min(ng, r0 + base) - r0;
            }
        }
        c.counters["ks_row_sharded"] = shard ? 1.0 : 0.0;
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
        qa.alloc((size_t)ng * b);      // the rank-revealing fallback QR runs on the gathered block
        qq.alloc((size_t)ng);
        if (shard) {
            Xfull.alloc((size_t)ng * b); Zfull.alloc((size_t)ng * b); Qfull.alloc((size_t)ng * b);
            const int gc = std::max(b, 128);           // widest gather: blocks of b columns, U in chunks of 128
            gpad.alloc((size_t)base * gc); grecv.alloc((size_t)base * gc * c.world);
            ctmp.alloc((size_t)ncv * kMaxB);
            gram1.alloc(256); gram2.alloc(256);
        }
        drank.alloc(1);
        dinfo.alloc(1);
        // deferred QR rank check (ks_defer_rank, default on): the rank of block step j is read from pinned memory only after
        // the operator of step j + 1 has been queued, so the host round trip hides behind ~0.4 ms of GPU work instead
        // of idling the device 41 (c2) / 1200 (c3) times per run; the next operator output goes to the other F buffer so that a
        // rank-deficient block (rare) can still be re-factorised by the fallback QR and the operator redone
        defer_rank = c.opt("ks_defer_rank", 1) != 0;
        force_fallback = c.opt("ks_force_qr_fallback", 0);
        F2.alloc((size_t)n * b);
        // the fast QR writes its rank straight into host-mapped memory: no D2H copy sits in the stream between the QR and
        // the next operator application, so nothing waits on the copy engine (which a background download of B may be
        // keeping busy: isle_cuda_download_B_begin)
        ISLE_CUDA_CHECK(cudaHostAlloc((void **)&h_rank, sizeof(int), cudaHostAllocMapped));
        ISLE_CUDA_CHECK(cudaHostGetDevicePointer((void **)&h_rank_dev, h_rank, 0));
        ISLE_CUDA_CHECK(cudaEventCreateWithFlags(&ev_rank, cudaEventDisableTiming));
        ISLE_CUDA_CHECK(cudaMemsetAsync(V.p, 0, V.bytes(), c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(H.p, 0, H.bytes(), c.stream));
        ISLE_CUSOLVER_CHECK(cusolverDnSsyevd_bufferSize(c.cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_UPPER,
                                                        m, Tm.p, m, Wev.p, &lwork));
        work.alloc((size_t)lwork);
        int per_sm = 0;
        ISLE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, mgs_qr64_kernel, 256, 0));
        ISLE_REQUIRE(per_sm >= 1, ISLE_ERR_CUDA, "mgs_qr64_kernel cannot be made resident");
        int want = (int)((ng + 255) / 256);
        qr_grid = std::max(1, std::min(want, c.num_sms));   // one CTA per SM at most
        qpart.alloc((size_t)2 * qr_grid * kMaxB);
        // CholeskyQR2 fast path: as many co-resident CTAs as the 16-row tiles can use
        ISLE_CUDA_CHECK(cudaFuncSetAttribute(cholqr2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholDynSmem));
        ISLE_CUDA_CHECK(cudaFuncSetAttribute(gram_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCholDynSmem));
        ISLE_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cholqr2_kernel, 256, kCholDynSmem));
        ISLE_REQUIRE(per_sm >= 1, ISLE_ERR_CUDA, "cholqr2_kernel cannot be made resident");
        chol_grid = std::max(1, std::min((int)((n + kCholTile - 1) / kCholTile), c.num_sms));   // every CTA re-sums all partials: keep them few
        cpart.alloc((size_t)2 * chol_grid * 256);
        fast_qr = c.opt("ks_fast_qr", 1) != 0;
        custom_orth = c.opt("ks_custom_orth", 1) != 0;
        panel_v2 = c.opt("ks_panel_v2", 1) != 0;
        ps.init(n, ncv, panel_v2);
        // tensor-core panel products (panel_tc.cu) once the basis has at least this many columns
        panel_tc = c.opt("ks_panel_tc", 1) != 0 && PanelTc::usable(n) && b <= 16;
        panel_tc_min_rows = shard ? 0 : c.opt("ks_panel_tc_min_rows", 0);
        ISLE_REQUIRE(!shard || panel_tc, ISLE_ERR_ARG, "block_ks: the row-sharded basis needs the tensor-core panel engine");
        ptc.sharded = shard;
        if (panel_tc) ptc.init(c, n, ncv);
        gs_elide = c.opt("ks_gs_elide", 1) != 0;
        gs_passes = std::max(2, std::min(3, c.opt("ks_gs_passes", 3)));
    }

    float *Vcol(int j) { return V.p + (size_t)j * n; }
    float *Hat(int r, int col) { return H.p + (size_t)col * ncv + r; }

    void randu(float *dst, int cols)
    {
        randu_kernel<<<grid_for((size_t)n * cols, 256), 256, 0, c.stream>>>(dst, (size_t)n, (size_t)ng, (size_t)r0, cols, seed, ++rng_calls);
        count_launch(c);
    }

    // full(ng x cols, ld ng) <- all ranks' row slices loc(n x cols, ld n); cols <= max(b, 128)
    void gather(const float *loc, int cols, float *full)
    {
        pad_rows_kernel<<<grid_for((size_t)base * cols, 256), 256, 0, c.stream>>>(loc, (size_t)n, (size_t)base, cols, gpad.p);
        allgather_f32(c, gpad.p, grecv.p, (size_t)base * cols);
        place_rows_kernel<<<grid_for((size_t)ng * cols, 256), 256, 0, c.stream>>>(grecv.p, (size_t)ng, (size_t)base, c.world, cols, full);
        count_launch(c, 2);
    }
    void slice(const float *full, int cols, float *loc)
    {
        slice_rows_kernel<<<grid_for((size_t)n * cols, 256), 256, 0, c.stream>>>(full, (size_t)ng, (size_t)r0, (size_t)n, cols, loc);
        count_launch(c);
    }

    // Rank-revealing fallback for a row-sharded block: gather it, run the reference's MGS on the whole block on every rank
    // (deterministic kernel, identical result), keep this rank's rows of Q.
    int qr_gathered(const float *Fsrc, int cols, float *Qdst)
    {
        gather(Fsrc, cols, Xfull.p);
        ISLE_CUDA_CHECK(cudaMemsetAsync(Qfull.p, 0, (size_t)ng * cols * 4, c.stream));
        ISLE_CUDA_CHECK(cudaMemsetAsync(Rb.p, 0, Rb.bytes(), c.stream));
        QrParams p;
        p.F = Xfull.p; p.n = ng; p.b = cols; p.a = qa.p; p.q = qq.p; p.Q = Qfull.p; p.R = Rb.p;
        p.rank_out = drank.p; p.part = qpart.p;
        void *args[] = {&p};
        ISLE_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mgs_qr64_kernel, dim3(qr_grid), dim3(256), args, 0, c.stream));
        count_launch(c);
        slice(Qfull.p, cols, Qdst);
        int rank = -1;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&rank, drank.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return rank;
    }

    // CholeskyQR2 fast path, launch only: Q -> Qdst, R -> Rb, the rank (or -1: bad pivot, the fallback must run) travels to
    // pinned memory behind an event.  Rb must have been zeroed by the caller.
    void qr_fast_launch(const float *Fsrc, int cols, float *Qdst)
    {
        CholQrParams q;
        q.F = Fsrc; q.n = n; q.b = cols; q.Q = Qdst; q.R = Rb.p; q.rank_out = h_rank_dev; q.part = cpart.p;
        if (shard) {
            gram_rows_kernel<<<chol_grid, 256, kCholDynSmem, c.stream>>>(q, nullptr);
            sum_partials_kernel<<<1, 256, 0, c.stream>>>(cpart.p, (unsigned)chol_grid, gram1.p);
            allreduce_sum_f64(c, gram1.p, 256);
            gram_rows_kernel<<<chol_grid, 256, kCholDynSmem, c.stream>>>(q, gram1.p);
            sum_partials_kernel<<<1, 256, 0, c.stream>>>(cpart.p, (unsigned)chol_grid, gram2.p);
            allreduce_sum_f64(c, gram2.p, 256);
            cholqr2_finish_kernel<<<chol_grid, 256, 0, c.stream>>>(q, gram1.p, gram2.p);
            count_launch(c, 5);
        } else {
            void *qargs[] = {&q};
            ISLE_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)cholqr2_kernel, dim3(chol_grid), dim3(256), qargs,
                                                        kCholDynSmem, c.stream));
            count_launch(c);
        }
        ISLE_CUDA_CHECK(cudaEventRecord(ev_rank, c.stream));
    }
    int qr_fast_result()
    {
        ISLE_CUDA_CHECK(cudaEventSynchronize(ev_rank));
        // test hook (ks_force_qr_fallback = N): every N-th fast QR is treated as a bad pivot, so the fallback / redo path runs
        if (force_fallback > 0 && ++fast_qr_calls % force_fallback == 0) return -1;
        return *h_rank;
    }
    // the reference's modified Gram-Schmidt (rank revealing), on the gathered block when the basis is row-sharded
    int qr_slow(const float *Fsrc, int cols, float *Qdst)
    {
        if (shard) return qr_gathered(Fsrc, cols, Qdst);
        ISLE_CUDA_CHECK(cudaMemsetAsync(Rb.p, 0, Rb.bytes(), c.stream));
        QrParams p;
        p.F = Fsrc; p.n = n; p.b = cols; p.a = qa.p; p.q = qq.p; p.Q = Qdst; p.R = Rb.p;
        p.rank_out = drank.p; p.part = qpart.p;
        void *args[] = {&p};
        ISLE_CUDA_CHECK(cudaLaunchCooperativeKernel((void *)mgs_qr64_kernel, dim3(qr_grid), dim3(256), args, 0, c.stream));
        count_launch(c);
        int rank = -1;
        ISLE_CUDA_CHECK(cudaMemcpyAsync(&rank, drank.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
        return rank;
    }

    // [Q,R] = mgs_qr64(Fsrc); Q -> Qdst (ld n); R -> Rb (b x b, zero padded).  Returns rank.
    int qr(const float *Fsrc, int cols, float *Qdst)
    {
        ISLE_CUDA_CHECK(cudaMemsetAsync(Rb.p, 0, Rb.bytes(), c.stream));
        StatScope s(c, "ks_qr");
        if (fast_qr) {
            qr_fast_launch(Fsrc, cols, Qdst);
            const int rank = qr_fast_result();
            if (rank >= 0) return rank;
            c.counters["ks_qr_fallbacks"] += 1.0;
        }
        return qr_slow(Fsrc, cols, Qdst);
    }
    // NOTE: for cols < b the R rows use leading dimension `cols`; callers that need R pass cols == b.

    void gemm_tn(int rows, int cols, const float *A, const float *Bm, float *Cm, int ldc)
    {   // Cm(rows x cols) = A(n x rows)^T Bm(n x cols); summed over the ranks' row slices when the basis is sharded
        const float one = 1.f, zero = 0.f;
        if (shard) {
            ISLE_CUBLAS_CHECK(cublasSgemm(c.cublas, CUBLAS_OP_T, CUBLAS_OP_N, rows, cols, (int)n, &one, A, (int)n, Bm,
                                          (int)n, &zero, ctmp.p, rows));
            allreduce_sum_f32(c, ctmp.p, (size_t)rows * cols);
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Cm, (size_t)ldc * 4, ctmp.p, (size_t)rows * 4, (size_t)rows * 4, cols,
                                              cudaMemcpyDeviceToDevice, c.stream));
            count_launch(c);
            return;
        }
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
        if (shard) {     // the operator works on whole vectors: gather the block, apply (all-reduce inside), keep my rows
            gather(X, b, Xfull.p);
            spsptr_multiply_dev(c, b, Xfull.p, Zfull.p);
            slice(Zfull.p, b, Z);
            return;
        }
        spsptr_multiply_dev(c, b, X, Z);
    }

    // U(ng x k, column-major) <- the first k columns of the (possibly row-sharded) basis
    void export_U(float *U)
    {
        if (!shard) {
            ISLE_CUDA_CHECK(cudaMemcpyAsync(U, V.p, (size_t)ng * k * 4, cudaMemcpyDeviceToDevice, c.stream));
            return;
        }
        for (int j0 = 0; j0 < k; j0 += 128) {
            const int w = std::min(128, k - j0);
            gather(Vcol(j0), w, U + (size_t)j0 * ng);
        }
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
    // A deferred rank check comes due: the fast QR of the block step that wrote Vcol(rows) from Fsrc.  true = full rank,
    // nothing to do; false = the block was re-factorised by the rank-revealing QR and / or refilled (Vcol(rows) and the R
    // block of H changed: whatever was computed from them since must be redone).
    bool resolve_rank(int rows, int cols, const float *Fsrc)
    {
        int rk = qr_fast_result();
        if (rk == b) return true;
        if (rk < 0) {
            c.counters["ks_qr_fallbacks"] += 1.0;
            StatScope s(c, "ks_qr");
            rk = qr_slow(Fsrc, b, Vcol(rows));
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(rows, cols), (size_t)ncv * 4, Rb.p, (size_t)b * 4, (size_t)b * 4, b,
                                              cudaMemcpyDeviceToDevice, c.stream));
        }
        if (rk < b) {
            ISLE_CUDA_CHECK(cudaMemsetAsync(Vcol(rows + rk), 0, (size_t)(b - rk) * n * 4, c.stream));
            refill(rows + rk, rows + b);
        }
        return false;
    }

    void expand()
    {
        const bool defer = defer_rank && fast_qr;
        bool pending = false;
        int p_rows = 0, p_cols = 0, step = 0;
        const float *p_F = nullptr;
        while (H_rows < ncv) {
            const int rows = H_rows, cols = H_cols;
            float *Fc = (defer && (step++ & 1)) ? F2.p : F.p;
            op(Vcol(cols), Fc);                                           // F = A V_k
            if (pending) {      // the rank of the block this operator was just applied to, read behind the queued work
                pending = false;
                if (!resolve_rank(p_rows, p_cols, p_F)) op(Vcol(cols), Fc);
            }
            {
                StatScope s(c, "ks_orth", 2.0 * gs_passes * (double)n * rows * 4.0, 4.0 * gs_passes * (double)n * rows * b);
                float *Hk = Hat(0, cols);
                if (custom_orth && panel_tc && rows >= panel_tc_min_rows) {
                    // The reference always runs three passes (restarted_block_ks.h:83-90).  When the second pass only
                    // moved coefficients below 1e-4 of the first's, what is left along W is below fp32 resolution
                    // of H and the third pass is elided on the device (no host round trip; ks_gs_elide = 0 keeps it).
                    ptc.begin_step(c);
                    ptc.split_F(c, Fc, b);
                    for (int pass = 0; pass < gs_passes; ++pass) {
                        ptc.wtf(c, V.p, rows, b, C.p, ncv, Hk, ncv, pass == 0, gs_elide ? pass : 0);
                        ptc.fsub(c, V.p, rows, b, Fc, gs_elide ? pass : 0);
                        if (gs_elide && pass == 1 && gs_passes > 2) ptc.decide_elision(c, 1e-4f);
                    }
                } else if (custom_orth) {
                    // Hk = W^T F ; F -= W Hk ; then (gs_passes - 1) x { Ck = W^T F ; F -= W Ck ; Hk += Ck }
                    for (int pass = 0; pass < gs_passes; ++pass) {
                        ps.wtf(V.p, rows, b, Fc, C.p, ncv, Hk, pass == 0);
                        ps.fsub(V.p, rows, b, C.p, ncv, Fc);
                    }
                } else {
                    gemm_tn(rows, b, V.p, Fc, Hk, ncv);                      // Hk = W^T F
                    gemm_sub(rows, b, V.p, Hk, ncv, Fc);                     // F -= W Hk
                    for (int pass = 0; pass + 1 < gs_passes; ++pass) {
                        gemm_tn(rows, b, V.p, Fc, C.p, ncv);                 // Ck = W^T F
                        gemm_sub(rows, b, V.p, C.p, ncv, Fc);                // F -= W Ck
                        add_block_kernel<<<(rows * b + 255) / 256, 256, 0, c.stream>>>(Hk, ncv, C.p, ncv, rows, b);
                        count_launch(c);
                    }
                }
                // new b rows of H are zero left of the R block
                zero_block_kernel<<<(b * (cols + b) + 255) / 256, 256, 0, c.stream>>>(Hat(rows, 0), ncv, b, cols + b);
                count_launch(c);
            }
            if (defer) {
                ISLE_CUDA_CHECK(cudaMemsetAsync(Rb.p, 0, Rb.bytes(), c.stream));
                {
                    StatScope s(c, "ks_qr");
                    qr_fast_launch(Fc, b, Vcol(rows));
                }
                ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(rows, cols), (size_t)ncv * 4, Rb.p, (size_t)b * 4, (size_t)b * 4, b,
                                                  cudaMemcpyDeviceToDevice, c.stream));
                pending = true;
                p_rows = rows; p_cols = cols; p_F = Fc;
                H_rows += b;
                H_cols += b;
                continue;
            }
            const int rk = qr(Fc, b, Vcol(rows));
            ISLE_CUDA_CHECK(cudaMemcpy2DAsync(Hat(rows, cols), (size_t)ncv * 4, Rb.p, (size_t)b * 4, (size_t)b * 4, b,
                                              cudaMemcpyDeviceToDevice, c.stream));
            H_rows += b;
            H_cols += b;
            if (rk < b) {
                ISLE_CUDA_CHECK(cudaMemsetAsync(Vcol(rows + rk), 0, (size_t)(b - rk) * n * 4, c.stream));
                refill(rows + rk, rows + b);
            }
        }
        if (pending) resolve_rank(p_rows, p_cols, p_F);
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
        read_small(c, &info, dinfo.p, sizeof(int));
        ISLE_REQUIRE(info == 0, ISLE_ERR_CUDA, "evd(H) failed");   // restarted_block_ks.h:156-157
        reverse_top_kernel<<<(mm * kk + 255) / 256, 256, 0, c.stream>>>(Tm.p, Wev.p, mm, kk, S.p, theta.p);
        count_launch(c);
        // V_mid <- V[:, nconv:m) S ; new starts V[:, m:m+b) move to [k, k+b)
        gemm_3xtf32(c, (int)n, kk, mm, Vcol(nconv), (int)n, S.p, mm, Vtmp.p, (int)n);     // tensor cores above ~20 GFLOP
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
        // replicated ranks must take identical restart decisions: rank 0's numbers win
        if (c.world > 1) bcast_f32(c, pack.p, h.size(), 0);
        read_small(c, h.data(), pack.p, h.size() * 4);
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
    build_csr(c, b <= 10 && c.opt("spmm_bfp", 1) != 0);
    KsState ks(c, k, b, seed);
    ks.init();
    int nconv = 0, n_restarts = 0;
    std::vector<float> evs, norms;
    c.counters["ks_unconverged"] = 0.0;
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
    if (n_restarts == max_restarts) {
        // restarted_block_ks.h:302-315: after the last expand() the reference reads H.tail_rows(blk) of the EXPANDED
        // ncv x 2k matrix.  Those rows are zero left of the final R block (columns >= 2k - b >= k), so its first
        // column with norm >= tol is >= k and nconv is clamped to nev: the reference carries on with the Ritz pairs of
        // the last truncation (which expand() leaves untouched in V[:, 0:k) and diag(H)[0:k]) and its assert
        // nconv == num_topics (src/sparseMatrix.cpp:1207) passes.  Same here; the number of pairs whose true residual is
        // still above tol is reported through the counter ks_unconverged instead of being lost.
        ks.residuals(evs, norms);
        int bad = 0;
        for (int j = 0; j < k; ++j)
            if (!(norms[j] / evs[j] < tol)) ++bad;
        c.counters["ks_unconverged"] = bad;
        nconv = k;
    }
    nconv = std::min(nconv, k);
    c.counters["ks_restarts"] = n_restarts;
    c.counters["ks_gs_elided"] = ks.panel_tc ? (double)ks.ptc.elided(c) : 0.0;
    c.counters["ks_nconv"] = nconv;

    // src/sparseMatrix.cpp:1209-1214: eigenvalues = diag(H)[0:k], eigenvectors = V[:, 0:k)
    if (evs.empty()) ks.residuals(evs, norms);
    if (evalues_out) std::copy(evs.begin(), evs.begin() + k, evalues_out);
    c.k = k64;
    c.U.alloc((size_t)c.V * k);
    ks.export_U(c.U.p);
    if (U_out)
        ISLE_CUDA_CHECK(cudaMemcpyAsync(U_out, c.U.p, (size_t)c.V * k * 4, cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    c.have_U = true;
    c.have_P = false;
    if (nconv_out) *nconv_out = nconv;
    if (nconv != k) throw Error(ISLE_ERR_NOCONV, "block_ks: only " + std::to_string(nconv) + " of " + std::to_string(k) + " eigenpairs converged");
}

// Harness: one C = W^T F, F -= W C pair on caller data with a chosen engine (0 = fp32 FMA scalar loads,
// 1 = fp32 FMA vector loads, 2 = tcgen05 split TF32), so the engines can be checked in isolation.
void panel_products(Ctx &c, int64_t n, int rows, int b, const float *W_host, float *F_host_inout, float *C_host_out, int engine)
{
    ISLE_REQUIRE(n >= 1 && rows >= 1 && b >= 1 && b <= kMaxB && W_host && F_host_inout && C_host_out, ISLE_ERR_ARG,
                 "panel_products: bad arguments");
    DevBuf<float> W((size_t)n * rows), F((size_t)n * b), C((size_t)rows * b);
    ISLE_CUDA_CHECK(cudaMemcpyAsync(W.p, W_host, W.bytes(), cudaMemcpyHostToDevice, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(F.p, F_host_inout, F.bytes(), cudaMemcpyHostToDevice, c.stream));
    if (engine == 2) {
        ISLE_REQUIRE(PanelTc::usable(n), ISLE_ERR_ARG, "panel_products: the tensor-core engine needs n % 4 == 0");
        PanelTc ptc;
        ptc.init(c, n, rows);
        ptc.split_F(c, F.p, b);
        ptc.wtf(c, W.p, rows, b, C.p, rows, nullptr, 0, true);
        ptc.fsub(c, W.p, rows, b, F.p);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));     // ptc's buffers are released on return
    } else {
        PanelSimt ps(c);
        ps.init(n, rows, engine == 1);
        ps.wtf(W.p, rows, b, F.p, C.p, rows, nullptr, true);
        ps.fsub(W.p, rows, b, C.p, rows, F.p);
        ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
    }
    ISLE_CUDA_CHECK(cudaMemcpyAsync(F_host_inout, F.p, F.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaMemcpyAsync(C_host_out, C.p, C.bytes(), cudaMemcpyDeviceToHost, c.stream));
    ISLE_CUDA_CHECK(cudaStreamSynchronize(c.stream));
}

}  // namespace isle
