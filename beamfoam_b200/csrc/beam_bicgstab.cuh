// beam_bicgstab.cuh — coupled bundles: block-sparse BiCGStab preconditioned by the per-beam block-Thomas.
//
// Inter-beam couplings turn the union of independent block-tridiagonal systems into one block-sparse system, the shape of
// the reference's multibeamFvBlockMatrix (fvBlockMatrices/multibeamFvBlockMatrix/multibeamFvBlockMatrix.C:298-470:
// per-beam d / l / u blocks plus 6x6 blocks between cells of different beams, block CSR), which the reference meant to
// solve with a preconditioned BiCGStab (numerics/blockLduMatrix/BlockLduSolvers/BlockBiCGStab/myBlockBiCGStabSolver.C:79-189).
// Neither is compiled in the reference and its contact model is not in its repository: the coupling here
// is SYNTHETIC (a linear spring between two cell centres, bf_set_couplings) and parity is pinned to the CPU oracle only.
//
//   cpl_add_couplings   the springs into the assembled rows: -k I on the translational diagonal blocks, the spring forces
//                       into the source (the +k I off-diagonal blocks stay implicit in the adjacency list)
//   cpl_factor          per beam: D'_i = D_i - L_i (D'_{i-1}^{-1} U_{i-1}); keeps D'^{-1}_i and uPrime_i = D'^{-1}_i U_i
//   cpl_precond         y = T^{-1} p for any right-hand side: z_i = D'^{-1}_i (p_i - L_i z_{i-1}),  y_i = z_i - uPrime_i y_{i+1}
//   cpl_amul            y = A x: the three block rows of a cell plus its coupling partners (gather over the adjacency
//                       list of the cell: no atomics, so a solve is reproducible bit for bit)
//   cpl_dots            up to three dot products and the six componentwise L1 sums of a vector, one block, fixed order
//
// The rows are the cell-parallel assembly's (beam_long.cuh: long_assemble, createBeamMesh ordering, [N][36]); the Newton
// update after the solve is the cell-parallel update (long_update_faces / long_update_cells).  This path serves validation-
// scale bundles: it is a correct GPU implementation of the coupled solve, not yet a tuned one.
#pragma once
#include "beam_long.cuh"

struct CoupledParams {
    int64_t N;
    int B;
    const int64_t* cellStart;  // [B+1]
    const int32_t* cellBeam;   // [N]
    const int64_t* adjPtr;     // [N+1] adjacency of the couplings (both directions)
    const int64_t* adjCol;     // partner cell
    const double* adjK;        // stiffness
    double *L, *D, *U, *R;     // the assembled rows (long_assemble)
    double *Dinv, *UP;         // the factors of the block-tridiagonal part
};

__global__ void cpl_add_couplings(const __grid_constant__ BeamParams p, const CoupledParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= q.N || q.adjPtr[g] == q.adjPtr[g + 1]) return;
    auto Wof = [&](int64_t c) {
        const int b = q.cellBeam[c];
        return ld3(p.W, tileRow((size_t)(b >> 5), p.nmax, (int)(c - q.cellStart[b]), 3, b & 31), TS);
    };
    const V3 Wg = Wof(g);
    double ksum = 0.0;
    V3 f = v3(0, 0, 0); // spring force on this cell
    for (int64_t e = q.adjPtr[g]; e < q.adjPtr[g + 1]; e++) {
        const double k = q.adjK[e];
        const V3 Wn = Wof(q.adjCol[e]);
        ksum += k;
        f = f + k * (Wn - Wg);
    }
    q.D[36 * g + 0] -= ksum; q.D[36 * g + 7] -= ksum; q.D[36 * g + 14] -= ksum;
    q.R[6 * g + 0] -= f.x; q.R[6 * g + 1] -= f.y; q.R[6 * g + 2] -= f.z;
}

// D (row-major 6x6, destroyed) -> its inverse by Gauss-Jordan with partial pivoting
__device__ inline void invert6(double* A, double* Ainv)
{
    for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) Ainv[6 * r + c] = r == c ? 1.0 : 0.0;
    luSolve6(A, Ainv, 6);
}

__global__ void cpl_factor(const CoupledParams q)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.B) return;
    double Dp[36], UPprev[36], tmp[36];
    for (int64_t g = q.cellStart[b]; g < q.cellStart[b + 1]; g++) {
        const bool firstCell = g == q.cellStart[b];
        for (int k = 0; k < 36; k++) Dp[k] = q.D[36 * g + k];
        if (!firstCell)
            for (int r = 0; r < 6; r++)
                for (int c = 0; c < 6; c++) {
                    double a = 0.0;
                    for (int k = 0; k < 6; k++) a += q.L[36 * g + 6 * r + k] * UPprev[6 * k + c];
                    Dp[6 * r + c] -= a;
                }
        invert6(Dp, tmp);
        for (int r = 0; r < 6; r++)
            for (int c = 0; c < 6; c++) {
                double a = 0.0;
                for (int k = 0; k < 6; k++) a += tmp[6 * r + k] * q.U[36 * g + 6 * k + c];
                UPprev[6 * r + c] = a;
            }
        for (int k = 0; k < 36; k++) { q.Dinv[36 * g + k] = tmp[k]; q.UP[36 * g + k] = UPprev[k]; }
    }
}

__global__ void cpl_precond(const CoupledParams q, const double* __restrict__ in, double* __restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= q.B) return;
    const int64_t g0 = q.cellStart[b], g1 = q.cellStart[b + 1];
    double z[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t g = g0; g < g1; g++) {
        double w[6];
        for (int r = 0; r < 6; r++) {
            double a = in[6 * g + r];
            if (g > g0) for (int k = 0; k < 6; k++) a -= q.L[36 * g + 6 * r + k] * z[k];
            w[r] = a;
        }
        for (int r = 0; r < 6; r++) {
            double a = 0.0;
            for (int k = 0; k < 6; k++) a += q.Dinv[36 * g + 6 * r + k] * w[k];
            z[r] = a;
        }
        for (int r = 0; r < 6; r++) out[6 * g + r] = z[r];
    }
    for (int64_t g = g1 - 2; g >= g0; g--) {
        double y[6];
        for (int r = 0; r < 6; r++) {
            double a = out[6 * g + r];
            for (int k = 0; k < 6; k++) a -= q.UP[36 * g + 6 * r + k] * out[6 * (g + 1) + k];
            y[r] = a;
        }
        for (int r = 0; r < 6; r++) out[6 * g + r] = y[r];
    }
}

__global__ void cpl_amul(const CoupledParams q, const double* __restrict__ x, double* __restrict__ y)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= q.N) return;
    const int b = q.cellBeam[g];
    const bool hasL = g > q.cellStart[b], hasU = g + 1 < q.cellStart[b + 1];
    double acc[6];
    for (int r = 0; r < 6; r++) {
        double a = 0.0;
        for (int k = 0; k < 6; k++) {
            a += q.D[36 * g + 6 * r + k] * x[6 * g + k];
            if (hasL) a += q.L[36 * g + 6 * r + k] * x[6 * (g - 1) + k];
            if (hasU) a += q.U[36 * g + 6 * r + k] * x[6 * (g + 1) + k];
        }
        acc[r] = a;
    }
    for (int64_t e = q.adjPtr[g]; e < q.adjPtr[g + 1]; e++) {
        const double k = q.adjK[e];
        const int64_t c = q.adjCol[e];
        acc[0] += k * x[6 * c]; acc[1] += k * x[6 * c + 1]; acc[2] += k * x[6 * c + 2];
    }
    for (int r = 0; r < 6; r++) y[6 * g + r] = acc[r];
}

// out[0..2] = a0.b0, a1.b1, a2.b2 (null pointers skipped); out[3..8] = componentwise L1 sums of l1 (6 components per cell)
__global__ void cpl_dots(const int64_t n, const double* __restrict__ a0, const double* __restrict__ b0, const double* __restrict__ a1,
                         const double* __restrict__ b1, const double* __restrict__ a2, const double* __restrict__ b2,
                         const double* __restrict__ l1, double* __restrict__ out)
{
    __shared__ double sm[9][256];
    double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        if (a0) v[0] += a0[i] * b0[i];
        if (a1) v[1] += a1[i] * b1[i];
        if (a2) v[2] += a2[i] * b2[i];
    }
    if (l1) // thread t sums component t % 6 when blockDim.x is a multiple of 6 ... keep it simple: stride 6 per component
        for (int64_t i = threadIdx.x; i < n / 6; i += blockDim.x)
            for (int k = 0; k < 6; k++) v[3 + k] += fabs(l1[6 * i + k]);
    for (int k = 0; k < 9; k++) sm[k][threadIdx.x] = v[k];
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o)
            for (int k = 0; k < 9; k++) sm[k][threadIdx.x] += sm[k][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x < 9) out[threadIdx.x] = sm[threadIdx.x][0];
}

// p = r + beta (p - omega v)
__global__ void cpl_update_p(const int64_t n, double* __restrict__ p, const double* __restrict__ r, const double* __restrict__ v,
                             const double beta, const double omega)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = r[i] + beta * p[i] - beta * omega * v[i];
}
// s = r - alpha v
__global__ void cpl_update_s(const int64_t n, double* __restrict__ s, const double* __restrict__ r, const double* __restrict__ v, const double alpha)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) s[i] = r[i] - alpha * v[i];
}
// x += alpha ph + omega sh ; r = s - omega t
__global__ void cpl_update_xr(const int64_t n, double* __restrict__ x, double* __restrict__ r, const double* __restrict__ ph,
                              const double* __restrict__ sh, const double* __restrict__ s, const double* __restrict__ t,
                              const double alpha, const double omega)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        x[i] = x[i] + alpha * ph[i] + omega * sh[i];
        r[i] = s[i] - omega * t[i];
    }
}
// the squared source norm of the coupled system into slot 0 of the partial sums (the other entries of the slot: zero)
__global__ void cpl_set_partial0(double* __restrict__ partial, const int nBlocks, const double* __restrict__ dots)
{
    for (int i = threadIdx.x; i < nBlocks; i += blockDim.x) partial[i] = i == 0 ? dots[0] : 0.0;
}
