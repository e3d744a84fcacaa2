// beam_coop.cuh — the Newton iteration for bundles BELOW ONE WAVE of beams: several lanes per beam.
//
// One thread per beam (beam_kernels.cuh) needs >= 37,888 beams to fill a B200 (148 SMs x 8 warps x 32 lanes); with fewer
// beams an iteration costs the latency of ONE thread walking its 200-cell recurrence (1.15 ms for 4096 as for 16,384
// beams).  Here a beam is owned by a GROUP of 4 lanes, a warp holds 8 beams, and the sweeps alternate between two lane
// mappings over batches of 4 consecutive cells:
//
//   forward   (a) assembly, one lane per CELL: lane (j, g) assembles cell i0+j of beam g -- face coefficients
//                 (updateEqnCoefficients, Evolve.C:673-753), own blocks and source (Matrix.C:55-475), end-patch closure
//                 (BC.C:64-1190), loads and inertia (Evolve.C:192-536) -- into a shared-memory record of the cell:
//                 the 6x6 own part of d_i, the source parts, and the five 3x3 tensors that define u_i, l_{i+1} and the
//                 neighbour part of d_{i+1}.  This part has no serial dependence and runs at full lane efficiency.
//             (b) elimination, one lane per COLUMN: the block-Thomas recurrence of the 4 cells, cell by cell; the 4 lanes
//                 of a group factor D'_i redundantly (2x2 block elimination over 3x3 adjugate inverses, pivoted 6x6 LU
//                 when a 3x3 determinant is small against its block) and lanes 0..2 each solve for two columns of
//                 uPrime_i, lane 3 for bPrime_i; every lane then forms its columns of the carry
//                 d_nei - l uPrime | srcNei - l bPrime.  The carry travels through shared memory (42 doubles per beam).
//   backward  back-substitution of 4 cells redundantly in the 4 lanes of the group (a 6x6 mat-vec per cell on the serial
//             chain), then one lane per cell: updateSolutionVariables (Evolve.C:757-923) and the patch evaluate()s.
//
// The per-cell critical path shrinks from a whole assembly + elimination (~5000 cycles) to the factorisation and one
// column solve (~700 cycles), and a bundle of B beams exposes 4B lanes.  State, scratch layout (uPrime/bPrime) and
// every per-cell / per-face formula are shared with the one-thread-per-beam kernels, so the two forward and the two
// backward kernels are interchangeable (tests/test_gpu_coop.py runs all four combinations).
#pragma once
#include "beam_kernels.cuh"

#define COOP_THREADS 32 // one warp = 8 beams per block: small blocks spread a small bundle over all SMs, and the
                        // shared memory of an SM divides into 7 warps (one more than with 2-warp blocks)
#define COOP_GROUP 4
#define COOP_BEAMS_PER_WARP 8
#define COOP_BEAMS_PER_BLOCK (COOP_THREADS / 32 * COOP_BEAMS_PER_WARP)
enum {
    K_DA = 0,   // 36: own contributions to d_i, row-major (face i+1 own part, inertia, end-patch closures)
    K_SO = 36,  // 6: own part of source_i (face i+1, loads, inertia, end patches)
    K_SN = 42,  // 6: neighbour part of source_{i+1} from face i+1: (eQ ; eM - eMQ)
    K_P = 48,   // 9: P = a CQW (full)
    K_QT = 57,  // 9: Qt = 0.5 CQTheta
    K_SU = 66,  // 9: S1 + S3 + S2, lower-right block of u_i     (S1 = a CMTheta, S3 = 0.5 CMQTheta, S2 = 0.5 CMTheta2)
    K_SL = 75,  // 9: S1 + S3 - S2, lower-right block of l_{i+1}
    K_SD = 84,  // 9: S3 - S2 - S1, lower-right block of the neighbour part of d_{i+1}
    K_ROWS = 93
};
// per warp: the records of one batch [K_ROWS][32 = cell slot j * 8 + beam g], the carry [7 columns][6][8 beams] and the
// neighbour source part of the last cell of the previous batch [6][8]
#define COOP_CARRY_OFF (K_ROWS * 32)
#define COOP_SNPREV_OFF (COOP_CARRY_OFF + 42 * 8)
// ... and the staging rows [ST_ROWS][32 lanes]: the face rows and W that the assembly of the NEXT batch reads, brought in by
// per-lane cp.async (LDGSTS, no register in flight) while the current batch is being eliminated
enum { ST_W = 0, ST_WN = 3, ST_LF = 6, ST_K = 10, ST_Q = 13, ST_M = 16, ST_ROWS = 19 };
#define COOP_STAGE_OFF (COOP_SNPREV_OFF + 6 * 8)
#define COOP_WARP_DOUBLES (COOP_STAGE_OFF + ST_ROWS * 32)
#define COOP_SMEM_BYTES ((COOP_THREADS / 32) * COOP_WARP_DOUBLES * sizeof(double))

// Pivoted fallback: D (row-major 36) x = cols, two right-hand sides per lane.
__device__ __noinline__ void coopSolvePivoted(const double* __restrict__ D36, double* __restrict__ cols12)
{
    double A[6][6], Bq[6][2];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) A[r][c] = D36[6 * r + c];
        Bq[r][0] = cols12[r];
        Bq[r][1] = cols12[6 + r];
    }
    solve6<2>(A, Bq);
#pragma unroll
    for (int r = 0; r < 6; r++) { cols12[r] = Bq[r][0]; cols12[6 + r] = Bq[r][1]; }
}

// ---- (a) assembly of one cell into its record (rec = warp slab + lane; row k at rec[k * 32]) -------------------------
// Returns nothing; the complete source of the cell is SO (this record) + SN (record of cell i-1).
template <int SCHEME>
DEV void coopAssembleCell(const BeamParams& p, const int b, const int i, const int n, double* __restrict__ rec,
                          const double* __restrict__ st, const double dZ, const double dcI, const V3 dR0, const V3 CQ, const V3 CM,
                          const double ARho, const V3 CI)
{
    const CellIn cell = loadCellIn<SCHEME>(p, b, i); // issued first: consumed after the face part
    double D[6][6], src[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        src[r] = 0.0;
#pragma unroll
        for (int c = 0; c < 6; c++) D[r][c] = 0.0;
    }
    const V3 Wi = v3(st[(ST_W + 0) * 32], st[(ST_W + 1) * 32], st[(ST_W + 2) * 32]);
    if (i == 0 || i == n - 1) { // end patches: assembleBoundaryConditions (BC.C:64-1190), additive in D and source
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        if (i == 0) patchClosure(p, 0, b, Wi, D36, s6);
        if (i == n - 1) patchClosure(p, 1, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            src[r] = s6[r];
#pragma unroll
            for (int c = 0; c < 6; c++) D[r][c] = D36[6 * r + c];
        }
    }
    if (i < n - 1) { // internal face i+1: the same folded form as fwdFacePart (beam_kernels.cuh)
        const V3 Wn = v3(st[(ST_WN + 0) * 32], st[(ST_WN + 1) * 32], st[(ST_WN + 2) * 32]);
        const Q4 qf = q4(st[(ST_LF + 0) * 32], st[(ST_LF + 1) * 32], st[(ST_LF + 2) * 32], st[(ST_LF + 3) * 32]);
        const V3 Kf = v3(st[(ST_K + 0) * 32], st[(ST_K + 1) * 32], st[(ST_K + 2) * 32]);
        const V3 Qf = v3(st[(ST_Q + 0) * 32], st[(ST_Q + 1) * 32], st[(ST_Q + 2) * 32]);
        const V3 Mf = v3(st[(ST_M + 0) * 32], st[(ST_M + 1) * 32], st[(ST_M + 2) * 32]);
        const double h = 0.5 * dZ;
        const V3 t = v3(dR0.x + dcI * (Wn.x - Wi.x), dR0.y + dcI * (Wn.y - Wi.y), dR0.z + dcI * (Wn.z - Wi.z));
        const V3 th = h * t;
        const M3 Lm = quatToMat(qf);
        const V3 hQ = 0.5 * Qf;
        const Sym3 Pm = conjDiagSym(Lm, dcI * CQ);                              // a*CQW :680
        const V3 eQ = mul(Lm, cmul(CQ, mulTSub(Lm, t, v3(1.0, 0.0, 0.0))));     // :682 with Gamma of :881-887
        const M3 Gh = mulSpinR(Pm, th);
        M3 Qt = Gh;                                                             // 0.5*CQTheta :686-693
        Qt.a[1] += hQ.z; Qt.a[2] -= hQ.y; Qt.a[3] -= hQ.z; Qt.a[5] += hQ.x; Qt.a[6] += hQ.y; Qt.a[7] -= hQ.x;
        const V3 eMQ = cross(th, eQ);                                           // :737-741
        const Sym3 S1 = conjDiagSym(Lm, dcI * CM);                              // a*CMTheta :699
        const V3 eM = mul(Lm, cmul(CM, Kf));                                    // :684
        M3 S3;                                                                  // 0.5*CMQTheta :718-735
        {
            const double d = dot(th, hQ);
#pragma unroll
            for (int r = 0; r < 3; r++) {
                const V3 y = cross(th, v3(Gh.a[r], Gh.a[3 + r], Gh.a[6 + r]));
                const double q = r == 0 ? hQ.x : r == 1 ? hQ.y : hQ.z;
                S3.a[3 * r] = fma(-q, th.x, y.x); S3.a[3 * r + 1] = fma(-q, th.y, y.y); S3.a[3 * r + 2] = fma(-q, th.z, y.z);
                S3.a[4 * r] += d;
            }
        }
        const V3 hM = 0.5 * Mf;                                                 // 0.5*CMTheta2 = -S(hM) :703
        const M3 Pf = full(Pm), S1f = full(S1);
        M3 S2 = m3zero();                                                       // -spin(hM)
        S2.a[1] = hM.z; S2.a[2] = -hM.y; S2.a[3] = -hM.z; S2.a[5] = hM.x; S2.a[6] = hM.y; S2.a[7] = -hM.x;
        const M3 Sm = S3 - S2;
        const M3 Eo = (S3 - S1f) + S2;                                          // -a*CMTheta + 0.5*CMQTheta + 0.5*CMTheta2
        addBlk(D, 0, 0, -1.0, Pf);
        addBlk(D, 0, 3, 1.0, Qt);
        addBlkT(D, 3, 0, 1.0, Qt);                                              // -a*CMQW = (0.5 CQTheta).T() :706-715
        addBlk(D, 3, 3, 1.0, Eo);
        subV(src, 0, eQ);
        subV(src, 3, eM + eMQ);
        const V3 sn2 = eM - eMQ;
        rec[(K_SN + 0) * 32] = eQ.x; rec[(K_SN + 1) * 32] = eQ.y; rec[(K_SN + 2) * 32] = eQ.z;
        rec[(K_SN + 3) * 32] = sn2.x; rec[(K_SN + 4) * 32] = sn2.y; rec[(K_SN + 5) * 32] = sn2.z;
        const M3 Su = (S1f + S3) + S2, Sl = S1f + Sm, Sd = Sm - S1f;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            rec[(K_P + k) * 32] = Pf.a[k];
            rec[(K_QT + k) * 32] = Qt.a[k];
            rec[(K_SU + k) * 32] = Su.a[k];
            rec[(K_SL + k) * 32] = Sl.a[k];
            rec[(K_SD + k) * 32] = Sd.a[k];
        }
    } else { // the last cell has no u, no l: zero tensors keep the column solve of the elimination finite
#pragma unroll
        for (int k = K_SN; k < K_ROWS; k++) rec[k * 32] = 0.0;
    }
    addLoadsAndInertia<SCHEME>(p, b, i, dZ, ARho, CI, cell, D, src);
#pragma unroll
    for (int r = 0; r < 6; r++) {
        rec[(K_SO + r) * 32] = src[r];
#pragma unroll
        for (int c = 0; c < 6; c++) rec[(K_DA + 6 * r + c) * 32] = D[r][c];
    }
}

// What the assembly of cell i of beam (T, lb) reads: W of the cell and of its right neighbour and the rows of face i+1 by
// cp.async into the lane's own staging column (nobody else touches it: no barrier needed, only cp.async.wait_all before
// the reads); the rows the inertia terms read are pulled into L2 and loaded directly at the top of the assembly.
template <int SCHEME>
DEV void coopStageCell(const BeamParams& p, const size_t T, const int lb, const int i, const int n, double* __restrict__ st)
{
    if (i >= n) return;
    const int Rc = p.nmax, Rf = p.nmax + 1;
    const size_t c3 = tileRow(T, Rc, i, 3, lb), c4 = tileRow(T, Rc, i, 4, lb);
#pragma unroll
    for (int k = 0; k < 3; k++) cpAsync8(st + (ST_W + k) * 32, p.W + c3 + k * TS);
    if (i + 1 < n) {
        const size_t w3 = tileRow(T, Rc, i + 1, 3, lb), f3 = tileRow(T, Rf, i + 1, 3, lb), f4 = tileRow(T, Rf, i + 1, 4, lb);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            cpAsync8(st + (ST_WN + k) * 32, p.W + w3 + k * TS);
            cpAsync8(st + (ST_K + k) * 32, p.K + f3 + k * TS);
            cpAsync8(st + (ST_Q + k) * 32, p.Q + f3 + k * TS);
            cpAsync8(st + (ST_M + k) * 32, p.M + f3 + k * TS);
        }
#pragma unroll
        for (int k = 0; k < 4; k++) cpAsync8(st + (ST_LF + k) * 32, p.Lf + f4 + k * TS);
    }
    if (SCHEME != BFK_STEADY) {
        pfRows<4>(p.Lam, c4, TS);
        pfRows<3>(p.Om, c3, TS);
        if (SCHEME == BFK_NEWMARK) {
            pfRows<3>(p.Ac, c3, TS); pfRows<3>(p.dOm, c3, TS);
            if (p.first) pfRows<3>(p.U, c3, TS);
        } else {
            pfRows<3>(p.U, c3, TS);
            if (!p.first) { pfRows<3>(p.Uold, c3, TS); pfRows<3>(p.Omold, c3, TS); }
        }
    }
}

// ---- (b) elimination of one cell by the 4 lanes of its group ---------------------------------------------------------
// rc = record column of the cell (warp slab + slot * 8 + g), cy = carry column of the beam (warp slab + COOP_CARRY_OFF + g).
// role 0..2: columns role and role+3 of [u_i]; role 3: the right-hand side.  Returns the lane's two columns of the new
// carry in nA, nB (the caller stores them between two warp barriers).
DEV void coopEliminateCell(const BeamParams& p, const size_t T, const int lb, const int i, const bool last, const int role,
                           const double* __restrict__ rc, const double* __restrict__ cy, double (&nA)[6], double (&nB)[6])
{
    M3 A, Bm, Cm, E;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) { // D'_i = carry (column-major) + own part (row-major)
            A.a[3 * r + c] = cy[(6 * c + r) * 8] + rc[(K_DA + 6 * r + c) * 32];
            Bm.a[3 * r + c] = cy[(6 * (c + 3) + r) * 8] + rc[(K_DA + 6 * r + c + 3) * 32];
            Cm.a[3 * r + c] = cy[(6 * c + r + 3) * 8] + rc[(K_DA + 6 * (r + 3) + c) * 32];
            E.a[3 * r + c] = cy[(6 * (c + 3) + r + 3) * 8] + rc[(K_DA + 6 * (r + 3) + c + 3) * 32];
        }
    const bool isRhs = role == 3;
    const int rr = isRhs ? 2 : role;
    // the lane's two right-hand-side columns: (a) u[:, role] = (P[:, role] ; -Qt[role, :]) or the source, (b) u[:, role+3]
    V3 a1, a2, b1, b2;
    {
        const V3 pc = v3(rc[(K_P + rr) * 32], rc[(K_P + 3 + rr) * 32], rc[(K_P + 6 + rr) * 32]);
        const V3 qr = v3(rc[(K_QT + 3 * rr) * 32], rc[(K_QT + 3 * rr + 1) * 32], rc[(K_QT + 3 * rr + 2) * 32]);
        const V3 s1 = v3(rc[(K_SO + 0) * 32] + cy[36 * 8], rc[(K_SO + 1) * 32] + cy[37 * 8], rc[(K_SO + 2) * 32] + cy[38 * 8]);
        const V3 s2 = v3(rc[(K_SO + 3) * 32] + cy[39 * 8], rc[(K_SO + 4) * 32] + cy[40 * 8], rc[(K_SO + 5) * 32] + cy[41 * 8]);
        a1 = isRhs ? s1 : pc;
        a2 = isRhs ? s2 : -qr;
        b1 = v3(rc[(K_QT + rr) * 32], rc[(K_QT + 3 + rr) * 32], rc[(K_QT + 6 + rr) * 32]);
        b2 = v3(rc[(K_SU + rr) * 32], rc[(K_SU + 3 + rr) * 32], rc[(K_SU + 6 + rr) * 32]);
    }
    double detA, detS, mA, mS;
    const M3 Ai = invDet(A, detA, mA);
    const M3 AiB = mul(Ai, Bm);
    const M3 S = subMulM(E, Cm, AiB);
    const M3 Si = invDet(S, detS, mS);
    V3 xa1, xa2, xb1, xb2;
    if (smallDet(detA, mA) || smallDet(detS, mS)) { // identical in the 4 lanes of the group (same D')
        double D36[36], cols[12];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                D36[6 * r + c] = A.a[3 * r + c]; D36[6 * r + c + 3] = Bm.a[3 * r + c];
                D36[6 * (r + 3) + c] = Cm.a[3 * r + c]; D36[6 * (r + 3) + c + 3] = E.a[3 * r + c];
            }
        cols[0] = a1.x; cols[1] = a1.y; cols[2] = a1.z; cols[3] = a2.x; cols[4] = a2.y; cols[5] = a2.z;
        cols[6] = b1.x; cols[7] = b1.y; cols[8] = b1.z; cols[9] = b2.x; cols[10] = b2.y; cols[11] = b2.z;
        coopSolvePivoted(D36, cols);
        xa1 = v3(cols[0], cols[1], cols[2]); xa2 = v3(cols[3], cols[4], cols[5]);
        xb1 = v3(cols[6], cols[7], cols[8]); xb2 = v3(cols[9], cols[10], cols[11]);
    } else {
        const V3 ya = mul(Ai, a1);
        xa2 = mul(Si, subMul(a2, Cm, ya));
        xa1 = subMul(ya, AiB, xa2);
        const V3 yb = mul(Ai, b1);
        xb2 = mul(Si, subMul(b2, Cm, yb));
        xb1 = subMul(yb, AiB, xb2);
    }
    // ---- uPrime columns / bPrime to the scratch (same layout as the one-thread-per-beam sweeps)
    if (isRhs) {
        const size_t b0 = tileRow(T, p.nmax, i, 6, lb);
        stS(p.bP + b0, xa1.x); stS(p.bP + b0 + TS, xa1.y); stS(p.bP + b0 + 2 * TS, xa1.z);
        stS(p.bP + b0 + 3 * TS, xa2.x); stS(p.bP + b0 + 4 * TS, xa2.y); stS(p.bP + b0 + 5 * TS, xa2.z);
    } else if (!last) {
        double* u = p.uP + tileRow(T, p.nmax, i, 36, lb);
        stS(u + (size_t)(role) * TS, xa1.x); stS(u + (size_t)(6 + role) * TS, xa1.y); stS(u + (size_t)(12 + role) * TS, xa1.z);
        stS(u + (size_t)(18 + role) * TS, xa2.x); stS(u + (size_t)(24 + role) * TS, xa2.y); stS(u + (size_t)(30 + role) * TS, xa2.z);
        stS(u + (size_t)(3 + role) * TS, xb1.x); stS(u + (size_t)(9 + role) * TS, xb1.y); stS(u + (size_t)(15 + role) * TS, xb1.z);
        stS(u + (size_t)(21 + role) * TS, xb2.x); stS(u + (size_t)(27 + role) * TS, xb2.y); stS(u + (size_t)(33 + role) * TS, xb2.z);
    }
    // ---- the lane's columns of the carry of cell i+1:  N - l X,
    //      l = [[P, -Qt], [Qt.T(), Sl]],  N = neighbour part of d_{i+1} = [[-P, -Qt], [-Qt.T(), Sd]] | srcNei
    {
        M3 Pf, Qt, Sl;
#pragma unroll
        for (int k = 0; k < 9; k++) { Pf.a[k] = rc[(K_P + k) * 32]; Qt.a[k] = rc[(K_QT + k) * 32]; Sl.a[k] = rc[(K_SL + k) * 32]; }
        const V3 sn1 = v3(rc[(K_SN + 0) * 32], rc[(K_SN + 1) * 32], rc[(K_SN + 2) * 32]);
        const V3 sn2 = v3(rc[(K_SN + 3) * 32], rc[(K_SN + 4) * 32], rc[(K_SN + 5) * 32]);
        const V3 sd = v3(rc[(K_SD + rr) * 32], rc[(K_SD + 3 + rr) * 32], rc[(K_SD + 6 + rr) * 32]);
        const V3 na1 = isRhs ? sn1 : -a1, na2 = isRhs ? sn2 : a2; // column role < 3 of N: (-P[:, c] ; -Qt[c, :])
        const V3 nb1 = -b1, nb2 = sd;                             // column role + 3 of N: (-Qt[:, k] ; Sd[:, k])
        const double xA1[3] = {xa1.x, xa1.y, xa1.z}, xA2[3] = {xa2.x, xa2.y, xa2.z};
        const double xB1[3] = {xb1.x, xb1.y, xb1.z}, xB2[3] = {xb2.x, xb2.y, xb2.z};
        const double nA1[3] = {na1.x, na1.y, na1.z}, nA2[3] = {na2.x, na2.y, na2.z};
        const double nB1[3] = {nb1.x, nb1.y, nb1.z}, nB2[3] = {nb2.x, nb2.y, nb2.z};
#pragma unroll
        for (int r = 0; r < 3; r++) {
            double ta = nA1[r], tb = nB1[r], ba = nA2[r], bb = nB2[r];
#pragma unroll
            for (int k = 0; k < 3; k++) { ta = fma(-Pf.a[3 * r + k], xA1[k], ta); tb = fma(-Pf.a[3 * r + k], xB1[k], tb); }
#pragma unroll
            for (int k = 0; k < 3; k++) { ta = fma(Qt.a[3 * r + k], xA2[k], ta); tb = fma(Qt.a[3 * r + k], xB2[k], tb); }
#pragma unroll
            for (int k = 0; k < 3; k++) { ba = fma(-Sl.a[3 * r + k], xA2[k], ba); bb = fma(-Sl.a[3 * r + k], xB2[k], bb); }
#pragma unroll
            for (int k = 0; k < 3; k++) { ba = fma(-Qt.a[3 * k + r], xA1[k], ba); bb = fma(-Qt.a[3 * k + r], xB1[k], bb); }
            nA[r] = ta; nA[3 + r] = ba; nB[r] = tb; nB[3 + r] = bb;
        }
    }
}

// Per-beam constants of the lane's beam, loaded once.
struct CoopBeam {
    int n;
    double dZ, dcI, ARho;
    V3 dR0, CQ, CM, CI;
};
template <int SCHEME>
DEV CoopBeam coopLoadBeam(const BeamParams& p, const int b, const bool live)
{
    CoopBeam k;
    const size_t s = p.Bpad;
    k.n = live ? p.nSeg[b] : 0;
    k.dZ = p.dZ[b];
    k.dcI = 1.0 / k.dZ;
    k.ARho = SCHEME != BFK_STEADY ? p.ARho[b] : 0.0;
    k.dR0 = v3(p.refL[b], p.refL[3 * s + b], p.refL[6 * s + b]);
    k.CQ = ld3(p.CQ, b, s);
    k.CM = ld3(p.CM, b, s);
    k.CI = SCHEME != BFK_STEADY ? ld3(p.CI, b, s) : v3(0, 0, 0);
    return k;
}

template <int SCHEME>
__global__ void __launch_bounds__(COOP_THREADS) beam_forward_coop(const BeamParams p)
{
    extern __shared__ __align__(16) unsigned char coop_smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane & 7, j = lane >> 3; // beam of the warp; cell slot (assembly) / column role (elimination)
    const int b = blockIdx.x * COOP_BEAMS_PER_BLOCK + w * COOP_BEAMS_PER_WARP + g;
    const bool live = b < p.B;
    double* warpSm = (double*)coop_smem + (size_t)w * COOP_WARP_DOUBLES;
    double* rec = warpSm + lane;
    double* cy = warpSm + COOP_CARRY_OFF + g;
    double* sp = warpSm + COOP_SNPREV_OFF + g;
    double* st = warpSm + COOP_STAGE_OFF + lane;
    const CoopBeam kb = coopLoadBeam<SCHEME>(p, b, live);
    const int n = kb.n;
    const int nW = __reduce_max_sync(0xffffffffu, n);
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31;
    double res = 0.0;
    // carry = 0, previous neighbour source = 0
    for (int k = j; k < 42; k += 4) cy[k * 8] = 0.0;
    if (j == 0) {
#pragma unroll
        for (int r = 0; r < 6; r++) sp[r * 8] = 0.0;
    }
    coopStageCell<SCHEME>(p, T, lb, j, n, st);
    __syncwarp();
    for (int i0 = 0; i0 < nW; i0 += COOP_GROUP) {
        // ---- (a) one lane per cell
        const int i = i0 + j;
        const bool mine = i < n;
        cpAsyncWaitAll();
        if (mine) coopAssembleCell<SCHEME>(p, b, i, n, rec, st, kb.dZ, kb.dcI, kb.dR0, kb.CQ, kb.CM, kb.ARho, kb.CI);
        coopStageCell<SCHEME>(p, T, lb, i + COOP_GROUP, n, st); // the next batch's rows arrive during the elimination
        __syncwarp();
        if (mine) { // ||source||^2 (EigenOF.C:253-282, x0 = 0): own part + neighbour part from face i
#pragma unroll
            for (int r = 0; r < 6; r++) {
                const double sPrev = j == 0 ? sp[r * 8] : warpSm[(K_SN + r) * 32 + lane - 8];
                const double s = rec[(K_SO + r) * 32] + sPrev;
                res = fma(s, s, res);
            }
        }
        __syncwarp();
        // ---- (b) one lane per column, cell by cell
#pragma unroll 1
        for (int jj = 0; jj < COOP_GROUP; jj++) {
            const int ic = i0 + jj;
            const bool active = ic < n;
            const bool last = ic == n - 1;
            double nA[6], nB[6];
            if (active) coopEliminateCell(p, T, lb, ic, last, j, warpSm + jj * 8 + g, cy, nA, nB);
            __syncwarp(); // every lane has read the old carry
            if (active && !last) {
                const int cA = j == 3 ? 6 : j;
#pragma unroll
                for (int r = 0; r < 6; r++) cy[(6 * cA + r) * 8] = nA[r];
                if (j < 3) {
#pragma unroll
                    for (int r = 0; r < 6; r++) cy[(6 * (j + 3) + r) * 8] = nB[r];
                }
            }
            __syncwarp();
        }
        // the neighbour source part of the batch's last cell outlives its record
        if (j == 3 && mine) {
#pragma unroll
            for (int r = 0; r < 6; r++) sp[r * 8] = rec[(K_SN + r) * 32];
        }
        __syncwarp();
    }
    double v[1] = {res};
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward: back-substitution redundantly in the group, update one lane per cell
// ---------------------------------------------------------------------------------------------------------------------
// The backward kernel takes no shared memory, so the whole L1 is a cache and a batch's rows (4 cells x 8 beams x ~77
// doubles = 20 KB per warp) fit it many times over: the rows of the NEXT batch are prefetched into L1 while this one is
// being updated, and the serial chain x_i = bPrime_i - uPrime_i x_{i+1} then reads L1 hits instead of exposing an L2
// round trip per cell (the kernel is launched after the forward sweep wrote the scratch, so L1 holds nothing stale).
DEV void pfL1(const double* __restrict__ p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
template <int NR>
DEV void pfRowsL1(const double* __restrict__ p, size_t i0)
{
#pragma unroll
    for (int k = 0; k < NR; k++) pfL1(p + i0 + (size_t)k * TS);
}
template <int SCHEME>
DEV void coopPrefetchBack(const BeamParams& p, const size_t T, const int lb, const int i, const int n)
{
    if (i < 0 || i >= n) return;
    const int Rc = p.nmax, Rf = p.nmax + 1;
    const size_t m3 = tileRow(T, Rc, i, 3, lb), g3 = tileRow(T, Rf, i + 1, 3, lb);
    pfRowsL1<36>(p.uP, tileRow(T, Rc, i, 36, lb));
    pfRowsL1<6>(p.bP, tileRow(T, Rc, i, 6, lb));
    pfRowsL1<3>(p.W, m3); pfRowsL1<3>(p.Th, m3);
    pfRowsL1<4>(p.Lam, tileRow(T, Rc, i, 4, lb));
    if (SCHEME == BFK_NEWMARK) {
        pfRowsL1<3>(p.Ac, m3); pfRowsL1<3>(p.dOm, m3);
        if (p.first) { pfRowsL1<3>(p.U, m3); pfRowsL1<3>(p.Om, m3); }
    }
    if (SCHEME == BFK_EULER) {
        pfRowsL1<3>(p.U, m3); pfRowsL1<3>(p.Om, m3);
        if (!p.first) { pfRowsL1<3>(p.Wold, m3); pfRowsL1<3>(p.Uold, m3); pfRowsL1<4>(p.Lamold, tileRow(T, Rc, i, 4, lb)); }
    }
    pfRowsL1<4>(p.Lf, tileRow(T, Rf, i + 1, 4, lb));
    pfRowsL1<3>(p.K, g3); pfRowsL1<3>(p.Q, g3); pfRowsL1<3>(p.M, g3);
}

template <int SCHEME>
__global__ void __launch_bounds__(COOP_THREADS) beam_backward_coop(const BeamParams p)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane & 7, j = lane >> 3;
    const int b = blockIdx.x * COOP_BEAMS_PER_BLOCK + w * COOP_BEAMS_PER_WARP + g;
    const bool live = b < p.B;
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31, Rc = p.nmax;
    const int n = live ? p.nSeg[b] : 0;
    const int nW = __reduce_max_sync(0xffffffffu, n);
    const double dZ = p.dZ[b];
    const double dcI = 1.0 / dZ, dcB = 2.0 / dZ, delta = 1.0 / dcB;
    const V3 CQ = ld3(p.CQ, b, p.Bpad), CM = ld3(p.CM, b, p.Bpad);
    V3 dR0, e0;
    {
        const M3 refL = ld9(p.refL, b, p.Bpad);
        dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
        e0 = mulT(refL, dR0);
    }
    double dx2 = 0.0, x2 = 0.0;
    double xn[6] = {0, 0, 0, 0, 0, 0};   // x of the cell right of the one being substituted
    V3 WoldCarry = v3(0, 0, 0);          // pre-update W of the first cell of the batch to the right
    const int iTop = nW > 0 ? ((nW - 1) / COOP_GROUP) * COOP_GROUP : -COOP_GROUP;
    coopPrefetchBack<SCHEME>(p, T, lb, iTop + j, n);
    for (int i0 = iTop; i0 >= 0; i0 -= COOP_GROUP) {
        coopPrefetchBack<SCHEME>(p, T, lb, i0 - COOP_GROUP + j, n);
        // ---- back substitution x_i = bPrime_i - uPrime_i & x_{i+1} of the batch, in every lane of the group
        double xm[6], xr[6]; // the lane's own cell (i0 + j) and its right neighbour
#pragma unroll
        for (int r = 0; r < 6; r++) { xm[r] = 0.0; xr[r] = xn[r]; }
#pragma unroll
        for (int jj = COOP_GROUP - 1; jj >= 0; jj--) {
            const int ic = i0 + jj;
            if (ic < n) {
                const size_t u0 = tileRow(T, Rc, ic, 36, lb), b0 = tileRow(T, Rc, ic, 6, lb);
                double x[6];
#pragma unroll
                for (int r = 0; r < 6; r++) x[r] = p.bP[b0 + (size_t)r * TS];
                if (ic < n - 1) {
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double acc = 0.0;
#pragma unroll
                        for (int c = 0; c < 6; c++) acc += p.uP[u0 + (size_t)(6 * r + c) * TS] * xn[c];
                        x[r] -= acc;
                    }
                }
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    xn[r] = x[r];
                    if (jj == j) xm[r] = x[r];
                    if (jj == j + 1) xr[r] = x[r];
                }
            }
        }
        // ---- one lane per cell: updateSolutionVariables (Evolve.C:757-923) and the patch evaluate()s
        const int i = i0 + j;
        const bool mine = i < n;
        const V3 DW = v3(xm[0], xm[1], xm[2]), DT = v3(xm[3], xm[4], xm[5]);
        V3 Wold = v3(0, 0, 0);
        if (mine) {
#pragma unroll
            for (int r = 0; r < 6; r++) dx2 += xm[r] * xm[r];
            updateCell<SCHEME>(p, T, lb, i, DW, DT, Wold, x2);
        }
        // the pre-update W of cell i+1 lives in the lane of the next slot (same beam: lane + 8), or came from the
        // previous batch for the last slot
        V3 WoldN;
        WoldN.x = __shfl_down_sync(0xffffffffu, Wold.x, 8);
        WoldN.y = __shfl_down_sync(0xffffffffu, Wold.y, 8);
        WoldN.z = __shfl_down_sync(0xffffffffu, Wold.z, 8);
        if (j == COOP_GROUP - 1) WoldN = WoldCarry;
        if (mine)
            updateFacesOfCell(p, b, i, n, CQ, CM, dR0, e0, dcI, dcB, delta, Wold, WoldN, DW, DT, v3(xr[0], xr[1], xr[2]),
                              v3(xr[3], xr[4], xr[5]));
        WoldCarry.x = __shfl_sync(0xffffffffu, Wold.x, g);
        WoldCarry.y = __shfl_sync(0xffffffffu, Wold.y, g);
        WoldCarry.z = __shfl_sync(0xffffffffu, Wold.z, g);
    }
    double v[2] = {dx2, x2};
    blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.x);
}
