// beam_long.cuh — the Newton iteration for FEW, LONG beams (BASELINE configs[1]: one beam of 10^4 cells).
//
// The fused sweeps of beam_kernels.cuh give one thread to a beam: perfect for a bundle, one thread busy for a
// single long beam.  Here every CELL gets a thread:
//   long_assemble      a1-a4 per cell: coefficients of its two faces, the blocks d_i, l_i (coupling to x_{i-1}),
//                      u_i (coupling to x_{i+1}) and source_i, exactly the block system the reference hands to
//                      BlockEigenSolverOF (EigenOF.C:60-175), in memory (this path is for small N);
//   long_pcr_load      copies the rows into the reduction buffers (second pass: with the residual b - A x as right-hand side)
//   long_pcr_step      a5 by block PARALLEL CYCLIC REDUCTION: in step s every row eliminates its couplings to rows
//                      i-s and i+s, ceil(log2 n) steps; 6x6 solves with partial pivoting inside the block;
//   long_pcr_finish    x_i (+)= d_i^{-1} source_i.  Cyclic reduction does not pivot ACROSS blocks and on a 10^4-cell
//                      beam (EA/EI ~ 10^3 per cell, cond ~ n^4) its solution carries ~1e-8 of noise, enough to cost the
//                      Newton loop an extra iteration against the reference's pivoted LU; one step of iterative
//                      refinement (a second reduction on the residual) brings it to the accuracy of a stable solve;
//   long_update_faces  a6/a7 for the faces a cell owns (needs the neighbour's OLD W, so it runs before ...)
//   long_update_cells  ... a6 for the cells, and the norms a8.
// State, layout and every per-cell / per-face formula are shared with the fused path (the same device functions).
#pragma once
#include "beam_kernels.cuh"

struct LongParams {
    int64_t N;                 // cells of this context
    const int32_t* cellBeam;   // [N] beam of a cell (host / createBeamMesh ordering)
    const int64_t* cellStart;  // [B+1]
    // block rows, [N][36] / [N][6]: the assembled system (kept for the residual) ...
    double *L0, *D0, *U0, *R0;
    // ... and two working copies (cyclic reduction ping-pongs between them)
    double *L[2], *D[2], *U[2], *R[2];
    int refine;                // 0: first solve, 1: refinement pass (right-hand side = residual, x accumulates)
    double* X;                 // [N][6] the solution
    int step;                  // PCR stride of this launch
    int src;                   // which copy holds the input rows
};

// Dense n x n solve with partial pivoting and nr right-hand sides, plain loops over local arrays (this path is
// launch-latency bound, not arithmetic bound).  A is destroyed, B becomes A^{-1} B.  Row-major.
__device__ inline void luSolve6(double* A, double* B, const int nr)
{
    for (int k = 0; k < 6; k++) {
        int piv = k;
        double mx = fabs(A[6 * k + k]);
        for (int r = k + 1; r < 6; r++)
            if (fabs(A[6 * r + k]) > mx) { mx = fabs(A[6 * r + k]); piv = r; }
        if (piv != k) {
            for (int c = 0; c < 6; c++) { const double t = A[6 * k + c]; A[6 * k + c] = A[6 * piv + c]; A[6 * piv + c] = t; }
            for (int c = 0; c < nr; c++) { const double t = B[nr * k + c]; B[nr * k + c] = B[nr * piv + c]; B[nr * piv + c] = t; }
        }
        const double ip = 1.0 / A[6 * k + k];
        for (int r = k + 1; r < 6; r++) {
            const double m = A[6 * r + k] * ip;
            if (m == 0.0) continue;
            for (int c = k + 1; c < 6; c++) A[6 * r + c] -= m * A[6 * k + c];
            for (int c = 0; c < nr; c++) B[nr * r + c] -= m * B[nr * k + c];
        }
    }
    for (int k = 5; k >= 0; k--) {
        const double ip = 1.0 / A[6 * k + k];
        for (int c = 0; c < nr; c++) {
            double s = B[nr * k + c];
            for (int j = k + 1; j < 6; j++) s -= A[6 * k + j] * B[nr * j + c];
            B[nr * k + c] = s * ip;
        }
    }
}

DEV void putBlk(double (&D)[6][6], int r0, int c0, double s, const M3& T)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) D[r0 + r][c0 + c] = s * T.a[3 * r + c];
}

template <int SCHEME>
__global__ void __launch_bounds__(128) long_assemble(const __grid_constant__ BeamParams p, const LongParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[1] = {0.0};
    if (g < q.N) {
        const int b = q.cellBeam[g];
        const int i = (int)(g - q.cellStart[b]), n = p.nSeg[b];
        const size_t T = (size_t)(b >> 5), sb = p.Bpad;
        const int lane = b & 31, Rc = p.nmax, Rf = p.nmax + 1;
        const double dZ = p.dZ[b], dcI = 1.0 / dZ;
        const V3 CQ = ld3(p.CQ, b, sb), CM = ld3(p.CM, b, sb);
        const M3 refL = ld9(p.refL, b, sb);
        const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]), e0 = mulT(refL, dR0);
        const V3 Wc = ld3(p.W, tileRow(T, Rc, i, 3, lane), TS);
        double D[6][6], Lb[6][6], Ub[6][6], src[6];
        for (int r = 0; r < 6; r++) {
            src[r] = 0.0;
            for (int c = 0; c < 6; c++) D[r][c] = Lb[r][c] = Ub[r][c] = 0.0;
        }
        // Matrix.C:84-470 with w = 0.5, a = deltaCoeffs: the face between cells (own, nei) adds to d[own], d[nei] and
        // source[own], source[nei], and defines u[f] (row own) and l[f] (row nei).  CMTheta2 = -spin(M).
        for (int side = 0; side < 2; side++) { // 0: the face on the left of the cell (this cell is nei), 1: on the right (own)
            const int j = i + side;
            if (j == 0 || j == n) { // end patch: assembleBoundaryConditions (BC.C:64-1190)
                double D36[36], s6[6];
                for (int r = 0; r < 6; r++) {
                    s6[r] = src[r];
                    for (int c = 0; c < 6; c++) D36[6 * r + c] = D[r][c];
                }
                patchClosure(p, side, b, Wc, D36, s6);
                for (int r = 0; r < 6; r++) {
                    src[r] = s6[r];
                    for (int c = 0; c < 6; c++) D[r][c] = D36[6 * r + c];
                }
                continue;
            }
            const V3 Wo = side ? ld3(p.W, tileRow(T, Rc, i + 1, 3, lane), TS) : ld3(p.W, tileRow(T, Rc, i - 1, 3, lane), TS);
            const V3 t = side ? dR0 + dcI * (Wo - Wc) : dR0 + dcI * (Wc - Wo);
            const size_t f3 = tileRow(T, Rf, j, 3, lane);
            const M3 Lm = quatToMat(ldq(p.Lf, tileRow(T, Rf, j, 4, lane), TS));
            const FaceCoef f = faceCoefficients(Lm, CQ, CM, gammaOf(Lm, e0, t), ld3(p.K, f3, TS), ld3(p.Q, f3, TS),
                                                ld3(p.M, f3, TS), t, dcI, false);
            const M3 S2 = -0.5 * spin(f.Mold);          // 0.5*CMTheta2
            const M3 Pm = dcI * f.CQW, Qt = 0.5 * f.CQTheta, Rm = dcI * f.CMQW, S1 = dcI * f.CMTheta, S3 = 0.5 * f.CMQTheta;
            if (side) { // own
                addBlk(D, 0, 0, -1.0, Pm); addBlk(D, 0, 3, 1.0, Qt);
                addBlk(D, 3, 0, -1.0, Rm); addBlk(D, 3, 3, 1.0, (S3 + S2) - S1);
                putBlk(Ub, 0, 0, 1.0, Pm); putBlk(Ub, 0, 3, 1.0, Qt);
                putBlk(Ub, 3, 0, 1.0, Rm); putBlk(Ub, 3, 3, 1.0, S1 + (S3 + S2));
                subV(src, 0, f.eQ); subV(src, 3, f.eM + f.eMQ);
            } else { // nei
                addBlk(D, 0, 0, -1.0, Pm); addBlk(D, 0, 3, -1.0, Qt);
                addBlk(D, 3, 0, 1.0, Rm); addBlk(D, 3, 3, 1.0, (S3 - S2) - S1);
                putBlk(Lb, 0, 0, 1.0, Pm); putBlk(Lb, 0, 3, -1.0, Qt);
                putBlk(Lb, 3, 0, -1.0, Rm); putBlk(Lb, 3, 3, 1.0, S1 + (S3 - S2));
                addV(src, 0, f.eQ); addV(src, 3, f.eM - f.eMQ);
            }
        }
        const double ARho = SCHEME != BFK_STEADY ? p.ARho[b] : 0.0;
        addLoadsAndInertia<SCHEME>(p, b, i, dZ, ARho, SCHEME != BFK_STEADY ? ld3(p.CI, b, p.Bpad) : v3(0, 0, 0), loadCellIn<SCHEME>(p, b, i), D, src);
        for (int r = 0; r < 6; r++) v[0] += src[r] * src[r];
        double *oD = q.D0 + 36 * g, *oL = q.L0 + 36 * g, *oU = q.U0 + 36 * g, *oR = q.R0 + 6 * g;
        for (int r = 0; r < 6; r++) {
            oR[r] = src[r];
            for (int c = 0; c < 6; c++) { oD[6 * r + c] = D[r][c]; oL[6 * r + c] = Lb[r][c]; oU[6 * r + c] = Ub[r][c]; }
        }
    }
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.x);
}

// Rows into working copy 0; right-hand side = source (first pass) or source - A x (refinement pass).
__global__ void __launch_bounds__(128) long_pcr_load(const __grid_constant__ BeamParams p, const LongParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= q.N) return;
    for (int k = 0; k < 36; k++) { q.L[0][36 * g + k] = q.L0[36 * g + k]; q.D[0][36 * g + k] = q.D0[36 * g + k]; q.U[0][36 * g + k] = q.U0[36 * g + k]; }
    double r[6];
    for (int k = 0; k < 6; k++) r[k] = q.R0[6 * g + k];
    if (q.refine) {
        const int b = q.cellBeam[g];
        const int i = (int)(g - q.cellStart[b]), n = p.nSeg[b];
        for (int k = 0; k < 6; k++) {
            double a = 0.0;
            for (int c = 0; c < 6; c++) {
                a += q.D0[36 * g + 6 * k + c] * q.X[6 * g + c];
                if (i > 0) a += q.L0[36 * g + 6 * k + c] * q.X[6 * (g - 1) + c];
                if (i < n - 1) a += q.U0[36 * g + 6 * k + c] * q.X[6 * (g + 1) + c];
            }
            r[k] -= a;
        }
    }
    for (int k = 0; k < 6; k++) q.R[0][6 * g + k] = r[k];
}

// One step of block parallel cyclic reduction with stride q.step.  Row i:  L_i x_{i-s} + D_i x_i + U_i x_{i+s} = R_i.
// Eliminating x_{i-s} with row i-s and x_{i+s} with row i+s leaves a row that couples x_i to x_{i-2s}, x_{i+2s}:
//     D_i <- D_i - L_i Z^-_U - U_i Z^+_L,   L_i <- -L_i Z^-_L,   U_i <- -U_i Z^+_U,   R_i <- R_i - L_i z^-_R - U_i z^+_R
// with Z^- = D_{i-s}^{-1} [L | U | R]_{i-s} and Z^+ the same of row i+s.
// SEVEN threads per row, one per COLUMN: thread c < 6 owns column c of the new L_i, D_i, U_i, thread 6 the right-hand side.
// Every thread factors the two neighbour diagonal blocks itself (pivoted 6x6, in registers) and solves only for its own
// columns, so the step is 7 N light threads instead of N heavy ones with 300 doubles of local arrays each (round 1: 70 us
// per step for 10^4 cells, latency-bound on local memory).
// Elimination of the coupling of row g to row h (SIDE 0: h = g - s through L_g, 1: h = g + s through U_g) for the thread's column c.
template <int SIDE>
DEV void pcrEliminateSide(const double* __restrict__ Din, const double* __restrict__ Lin, const double* __restrict__ Uin,
                          const double* __restrict__ Rin, const int64_t g, const int64_t h, const bool exists, const int c,
                          double (&colD)[6], double (&colL)[6], double (&colU)[6])
{
    if (!exists) return;
    double A[6][6], Z[6][2]; // Z = D_h^{-1} [column c of L_h | column c of U_h]   (thread 6: [R_h | 0])
    const int cc = c < 6 ? c : 0;
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int k = 0; k < 6; k++) A[r][k] = Din[36 * h + 6 * r + k];
        const double lv = Lin[36 * h + 6 * r + cc], uv = Uin[36 * h + 6 * r + cc], rv = Rin[6 * h + r];
        Z[r][0] = c < 6 ? lv : rv;
        Z[r][1] = c < 6 ? uv : 0.0;
    }
    solve6reg<2>(A, Z);
    const double* C = (SIDE ? Uin : Lin) + 36 * g; // the coupling of row g to row h
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double a0 = 0.0, a1 = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) { const double ck = C[6 * r + k]; a0 = fma(ck, Z[k][0], a0); a1 = fma(ck, Z[k][1], a1); }
        if (SIDE) { colD[r] -= a0; colU[r] = c < 6 ? -a1 : 0.0; }   // row g+s couples back to x_g through its L (thread 6: R_g -= C z_R)
        else { colD[r] -= c < 6 ? a1 : a0; colL[r] = c < 6 ? -a0 : 0.0; } // row g-s couples back to x_g through its U
    }
}

#define PCR_COLS 7
__global__ void __launch_bounds__(64, 8) long_pcr_step(const __grid_constant__ BeamParams p, const LongParams q)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t g = t / PCR_COLS;
    const int c = (int)(t - g * PCR_COLS);
    if (g >= q.N) return;
    const int b = q.cellBeam[g];
    const int i = (int)(g - q.cellStart[b]), n = p.nSeg[b], s = q.step, in = q.src, out = 1 - q.src;
    const double *Din = q.D[in], *Lin = q.L[in], *Uin = q.U[in], *Rin = q.R[in];
    double colD[6], colL[6], colU[6];
    if (c < 6) {
#pragma unroll
        for (int r = 0; r < 6; r++) { colD[r] = Din[36 * g + 6 * r + c]; colL[r] = 0.0; colU[r] = 0.0; }
    } else {
#pragma unroll
        for (int r = 0; r < 6; r++) { colD[r] = Rin[6 * g + r]; colL[r] = 0.0; colU[r] = 0.0; }
    }
    pcrEliminateSide<0>(Din, Lin, Uin, Rin, g, g - s, i - s >= 0, c, colD, colL, colU);
    pcrEliminateSide<1>(Din, Lin, Uin, Rin, g, g + s, i + s < n, c, colD, colL, colU);
    if (c < 6) {
#pragma unroll
        for (int r = 0; r < 6; r++) {
            q.D[out][36 * g + 6 * r + c] = colD[r];
            q.L[out][36 * g + 6 * r + c] = colL[r];
            q.U[out][36 * g + 6 * r + c] = colU[r];
        }
    } else {
#pragma unroll
        for (int r = 0; r < 6; r++) q.R[out][6 * g + r] = colD[r];
    }
}

__global__ void __launch_bounds__(128) long_pcr_finish(const LongParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= q.N) return;
    double A[36], x[6];
    for (int k = 0; k < 36; k++) A[k] = q.D[q.src][36 * g + k];
    for (int k = 0; k < 6; k++) x[k] = q.R[q.src][6 * g + k];
    luSolve6(A, x, 1);
    for (int k = 0; k < 6; k++) q.X[6 * g + k] = q.refine ? q.X[6 * g + k] + x[k] : x[k];
}

// Faces first (they read the neighbour cell's pre-update W) ...
__global__ void __launch_bounds__(128) long_update_faces(const __grid_constant__ BeamParams p, const LongParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= q.N) return;
    const int b = q.cellBeam[g];
    const int i = (int)(g - q.cellStart[b]), n = p.nSeg[b];
    const size_t T = (size_t)(b >> 5), sb = p.Bpad;
    const int lane = b & 31, Rc = p.nmax;
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ, delta = 1.0 / dcB;
    const V3 CQ = ld3(p.CQ, b, sb), CM = ld3(p.CM, b, sb);
    const M3 refL = ld9(p.refL, b, sb);
    const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]), e0 = mulT(refL, dR0);
    const double* x = q.X + 6 * g;
    const V3 DW = v3(x[0], x[1], x[2]), DT = v3(x[3], x[4], x[5]);
    V3 xWn = v3(0, 0, 0), xTn = v3(0, 0, 0), WoldN = v3(0, 0, 0);
    if (i < n - 1) {
        xWn = v3(x[6], x[7], x[8]); xTn = v3(x[9], x[10], x[11]);
        WoldN = ld3(p.W, tileRow(T, Rc, i + 1, 3, lane), TS);
    }
    updateFacesOfCell(p, b, i, n, CQ, CM, dR0, e0, dcI, dcB, delta, ld3(p.W, tileRow(T, Rc, i, 3, lane), TS), WoldN, DW, DT, xWn, xTn);
}
// ... then the cells, and the two solution norms.
template <int SCHEME>
__global__ void __launch_bounds__(128) long_update_cells(const __grid_constant__ BeamParams p, const LongParams q)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (g < q.N) {
        const int b = q.cellBeam[g];
        const int i = (int)(g - q.cellStart[b]);
        const double* x = q.X + 6 * g;
        for (int r = 0; r < 6; r++) v[0] += x[r] * x[r];
        V3 Wold;
        updateCell<SCHEME>(p, (size_t)(b >> 5), b & 31, i, v3(x[0], x[1], x[2]), v3(x[3], x[4], x[5]), Wold, v[1]);
    }
    blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.x);
}
