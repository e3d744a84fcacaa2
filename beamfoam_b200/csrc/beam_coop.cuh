// beam_coop.cuh — the Newton iteration for bundles BELOW ONE WAVE of beams.
//
// One thread per beam (beam_kernels.cuh) needs >= 37,888 beams to fill a B200 (148 SMs x 8 warps x 32 lanes); with fewer
// beams an iteration costs the latency of ONE thread walking its 200-cell recurrence through assembly AND elimination
// (~5000 cycles per cell: 1.15 ms for 4096 as for 16,384 beams).  Only the block-Thomas elimination is serial along a
// beam; everything else of the iteration is independent per cell.  These kernels split the two:
//
//   beam_forward_ws  WARP-SPECIALISED, one block = 16 beams = two warps that run concurrently on two SM sub-partitions:
//       assembler warp    one lane per CELL (2 consecutive cells x 16 beams per pass): face coefficients
//                         (updateEqnCoefficients, Evolve.C:673-753), own blocks and source (Matrix.C:55-475), end-patch
//                         closure (BC.C:64-1190), loads and inertia (Evolve.C:192-536) -> a shared-memory RECORD per cell:
//                         the 3x3 tensors that define u_i, l_{i+1} and the face parts of d_i, d_{i+1}, the inertia
//                         block and the source parts.  No serial dependence: full lane efficiency, runs ahead of ...
//       eliminator warp   two lanes per beam, one lane per group of COLUMNS: the block-Thomas recurrence, cell by cell.
//                         Both lanes factor D'_i (2x2 block elimination over 3x3 adjugate inverses; pivoted 6x6 LU when
//                         a 3x3 determinant is small against its block), lane 0 solves for columns 0..2 of uPrime_i and
//                         for bPrime_i, lane 1 for columns 3..5; each lane then forms its columns of the carry
//                         d_nei - l uPrime | srcNei - l bPrime, which travels through shared memory.
//       The records are double-buffered; the warps hand the buffers over with named barriers (bar.sync / bar.arrive:
//       full[2], empty[2]).  The serial chain of a beam shrinks to factorisation + column solves (~1200 cycles per cell).
//   beam_backsub     the serial part of the backward sweep only: x_i = bPrime_i - uPrime_i x_{i+1}, one thread per beam,
//                    uPrime/bPrime streamed through a 4-cell cp.async ring in shared memory (the loads do not depend on x,
//                    so the chain never waits for memory), x written to a tiled [cell][6][32] array.
//   beam_update_faces / beam_update_cells   updateSolutionVariables (Evolve.C:757-923) and the patch evaluate()s with one
//                    thread per CELL over the whole bundle: no serial dependence once x is known, so the latency of this
//                    (HBM-bound) part is hidden by occupancy, not by the length of a beam.  Faces first (they read the
//                    neighbour cell's pre-update W), then cells and the norms.
//
// State, scratch layout (uPrime/bPrime) and every per-cell / per-face formula are shared with the one-thread-per-beam
// kernels, so the two forward and the two backward implementations are interchangeable (tests/test_gpu_coop.py runs all
// four combinations against the oracle).
#pragma once
#include "beam_kernels.cuh"

#define WS_THREADS 64
#define WS_BEAMS 16           // beams per block
#define WS_BATCH 2            // cells per beam per assembler pass
// record rows (per cell; row k of column col at rec[k * WS_COLS + col], col = buffer * 32 + slot * 16 + beam)
enum {
    K_DI = 0,   // 10: inertia part of d_i: QRhoCoeff, then the 3x3 MRhoCoeff block (Evolve.C:459-535)
    K_SO = 10,  // 6: own part of source_i (face i+1, loads, inertia, end patches)
    K_SN = 16,  // 6: neighbour part of source_{i+1} from face i+1: (eQ ; eM - eMQ)
    K_P = 22,   // 9: P = a CQW
    K_QT = 31,  // 9: Qt = 0.5 CQTheta
    K_S1 = 40,  // 9: S1 = a CMTheta
    K_S3 = 49,  // 9: S3 = 0.5 CMQTheta
    K_HM = 58,  // 3: hM = 0.5 M of the face (0.5 CMTheta2 = -spin(hM))
    K_ROWS = 61,
    K_PD = K_SN // the LAST cell of a beam has no face i+1: rows K_SN .. K_SN+35 hold the right-patch closure (6x6, row-major)
};
#define WS_COLS 64
#define WS_CARRY_OFF (K_ROWS * WS_COLS)        // carry [7 columns][6 rows][16 beams]
#define WS_SMEM_DOUBLES (WS_CARRY_OFF + 42 * WS_BEAMS)
#define WS_SMEM_BYTES (WS_SMEM_DOUBLES * sizeof(double))
enum { BAR_FULL = 1, BAR_EMPTY = 3 }; // named barriers 1,2 / 3,4 (0 is __syncthreads)

DEV void namedSync(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(WS_THREADS) : "memory"); }
DEV void namedArrive(int id) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(WS_THREADS) : "memory"); }

// Pivoted fallback of the elimination: D (row-major 36) x = cols, four right-hand sides per lane.
__device__ __noinline__ void wsSolvePivoted(const double* __restrict__ D36, double* __restrict__ cols24)
{
    double A[6][6], Bq[6][4];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) A[r][c] = D36[6 * r + c];
#pragma unroll
        for (int s = 0; s < 4; s++) Bq[r][s] = cols24[6 * s + r];
    }
    solve6<4>(A, Bq);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int s = 0; s < 4; s++) cols24[6 * s + r] = Bq[r][s];
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

// L2 prefetch of what the assembly of cell i of beam (T, lb) reads
template <int SCHEME>
DEV void wsPrefetchCell(const BeamParams& p, const size_t T, const int lb, const int i, const int n)
{
    if (i >= n) return;
    const int Rc = p.nmax, Rf = p.nmax + 1;
    const size_t c3 = tileRow(T, Rc, i, 3, lb), c4 = tileRow(T, Rc, i, 4, lb);
    pfRows<3>(p.W, c3, TS);
    if (i + 1 < n) {
        const size_t f3 = tileRow(T, Rf, i + 1, 3, lb);
        pfRows<4>(p.Lf, tileRow(T, Rf, i + 1, 4, lb), TS);
        pfRows<3>(p.K, f3, TS); pfRows<3>(p.Q, f3, TS); pfRows<3>(p.M, f3, TS);
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

// ---- assembler: one cell into its record (rec = record base + column; row k at rec[k * WS_COLS]) -------------------------
// cyInit (cell 0 only): the carry of the beam, initialised with the left-patch closure.  Returns the own part of the
// source (K_SO) in so[]; the complete source of the cell is so + SN of cell i-1.
template <int SCHEME>
DEV void wsAssembleCell(const BeamParams& p, const int b, const int i, const int n, double* __restrict__ rec,
                        double* __restrict__ cyInit, const CoopBeam& kb, double (&so)[6])
{
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31, Rc = p.nmax, Rf = p.nmax + 1;
    const CellIn cell = loadCellIn<SCHEME>(p, b, i); // issued first: consumed after the face part
    double src[6];
#pragma unroll
    for (int r = 0; r < 6; r++) src[r] = 0.0;
    const V3 Wi = ld3(p.W, tileRow(T, Rc, i, 3, lb), TS);
    if (i == 0) { // left patch (BC.C:64-1190): its 6x6 part starts the carry, its source part joins the own source
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        patchClosure(p, 0, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            src[r] = s6[r];
            cyInit[(36 + r) * WS_BEAMS] = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) cyInit[(6 * c + r) * WS_BEAMS] = D36[6 * r + c]; // the carry is column-major
        }
    }
    if (i < n - 1) { // internal face i+1: the same folded form as fwdFacePart (beam_kernels.cuh)
        const size_t f3 = tileRow(T, Rf, i + 1, 3, lb);
        const V3 Wn = ld3(p.W, tileRow(T, Rc, i + 1, 3, lb), TS);
        const Q4 qf = ldq(p.Lf, tileRow(T, Rf, i + 1, 4, lb), TS);
        const V3 Kf = ld3(p.K, f3, TS), Qf = ld3(p.Q, f3, TS), Mf = ld3(p.M, f3, TS);
        const double h = 0.5 * kb.dZ, dcI = kb.dcI;
        const V3 t = v3(kb.dR0.x + dcI * (Wn.x - Wi.x), kb.dR0.y + dcI * (Wn.y - Wi.y), kb.dR0.z + dcI * (Wn.z - Wi.z));
        const V3 th = h * t;
        const M3 Lm = quatToMat(qf);
        const V3 hQ = 0.5 * Qf;
        const Sym3 Pm = conjDiagSym(Lm, dcI * kb.CQ);                           // a*CQW :680
        const V3 eQ = mul(Lm, cmul(kb.CQ, mulTSub(Lm, t, v3(1.0, 0.0, 0.0))));  // :682 with Gamma of :881-887
        const M3 Gh = mulSpinR(Pm, th);
        M3 Qt = Gh;                                                             // 0.5*CQTheta :686-693
        Qt.a[1] += hQ.z; Qt.a[2] -= hQ.y; Qt.a[3] -= hQ.z; Qt.a[5] += hQ.x; Qt.a[6] += hQ.y; Qt.a[7] -= hQ.x;
        const V3 eMQ = cross(th, eQ);                                           // :737-741
        const Sym3 S1 = conjDiagSym(Lm, dcI * kb.CM);                           // a*CMTheta :699
        const V3 eM = mul(Lm, cmul(kb.CM, Kf));                                 // :684
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
        subV(src, 0, eQ);
        subV(src, 3, eM + eMQ);
        const V3 sn2 = eM - eMQ;
        rec[(K_SN + 0) * WS_COLS] = eQ.x; rec[(K_SN + 1) * WS_COLS] = eQ.y; rec[(K_SN + 2) * WS_COLS] = eQ.z;
        rec[(K_SN + 3) * WS_COLS] = sn2.x; rec[(K_SN + 4) * WS_COLS] = sn2.y; rec[(K_SN + 5) * WS_COLS] = sn2.z;
#pragma unroll
        for (int k = 0; k < 9; k++) {
            rec[(K_P + k) * WS_COLS] = Pf.a[k];
            rec[(K_QT + k) * WS_COLS] = Qt.a[k];
            rec[(K_S1 + k) * WS_COLS] = S1f.a[k];
            rec[(K_S3 + k) * WS_COLS] = S3.a[k];
        }
        rec[(K_HM + 0) * WS_COLS] = hM.x; rec[(K_HM + 1) * WS_COLS] = hM.y; rec[(K_HM + 2) * WS_COLS] = hM.z;
    } else { // the last cell: right patch instead of a face; its 6x6 part takes the tensor rows of the record
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        patchClosure(p, 1, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) src[r] += s6[r];
#pragma unroll
        for (int k = 0; k < 36; k++) rec[(K_PD + k) * WS_COLS] = D36[k];
    }
    double D[6][6];
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) D[r][c] = 0.0;
    addLoadsAndInertia<SCHEME>(p, b, i, kb.dZ, kb.ARho, kb.CI, cell, D, src); // touches D[0..2][0..2] (diagonal) and D[3..5][3..5]
    rec[K_DI * WS_COLS] = D[0][0];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) rec[(K_DI + 1 + 3 * r + c) * WS_COLS] = D[3 + r][3 + c];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        rec[(K_SO + r) * WS_COLS] = src[r];
        so[r] = src[r];
    }
}

// ---- eliminator: one cell by the 2 lanes of its beam ---------------------------------------------------------------------
// rc = record column of the cell, cy = carry column of the beam (+ g).  role 0: columns 0..2 of u_i and the right-hand
// side; role 1: columns 3..5 (and an idle fourth slot).  Returns the lane's columns of the new carry in nc[slot][row].
DEV void wsEliminateCell(const BeamParams& p, const size_t T, const int lb, const int i, const bool last, const int role,
                         const double* __restrict__ rc, const double* __restrict__ cy, double (&nc)[4][6])
{
#define RC(k) rc[(k) * WS_COLS]
#define CY(c, r) cy[(6 * (c) + (r)) * WS_BEAMS]
    M3 Pf, Qt, S1, S3, S2;
    M3 A, Bm, Cm, E;
    const double qd = RC(K_DI);
    if (!last) {
#pragma unroll
        for (int k = 0; k < 9; k++) { Pf.a[k] = RC(K_P + k); Qt.a[k] = RC(K_QT + k); S1.a[k] = RC(K_S1 + k); S3.a[k] = RC(K_S3 + k); }
        const V3 hM = v3(RC(K_HM), RC(K_HM + 1), RC(K_HM + 2));
        S2 = m3zero(); // -spin(hM)
        S2.a[1] = hM.z; S2.a[2] = -hM.y; S2.a[3] = -hM.z; S2.a[5] = hM.x; S2.a[6] = hM.y; S2.a[7] = -hM.x;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { // D'_i = carry (column-major) + own part: [[-P + q I, Qt], [Qt.T(), S3 - S1 + S2 + inertia]]
                A.a[3 * r + c] = CY(c, r) - Pf.a[3 * r + c] + (r == c ? qd : 0.0);
                Bm.a[3 * r + c] = CY(c + 3, r) + Qt.a[3 * r + c];
                Cm.a[3 * r + c] = CY(c, r + 3) + Qt.a[3 * c + r];
                E.a[3 * r + c] = CY(c + 3, r + 3) + (((S3.a[3 * r + c] - S1.a[3 * r + c]) + S2.a[3 * r + c]) + RC(K_DI + 1 + 3 * r + c));
            }
    } else {
        Pf = Qt = S1 = S3 = S2 = m3zero();
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { // carry + right-patch closure + inertia
                A.a[3 * r + c] = CY(c, r) + RC(K_PD + 6 * r + c) + (r == c ? qd : 0.0);
                Bm.a[3 * r + c] = CY(c + 3, r) + RC(K_PD + 6 * r + c + 3);
                Cm.a[3 * r + c] = CY(c, r + 3) + RC(K_PD + 6 * (r + 3) + c);
                E.a[3 * r + c] = CY(c + 3, r + 3) + (RC(K_PD + 6 * (r + 3) + c + 3) + RC(K_DI + 1 + 3 * r + c));
            }
    }
    // ---- the lane's right-hand sides: slots 0..2 = u[:, 3 role + s], slot 3 = source - l bPrime (role 0 only)
    //      u = [[P, Qt], [-Qt.T(), S1 + S3 + S2]]
    double t1[4][3], t2[4][3];
#pragma unroll
    for (int s = 0; s < 3; s++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            t1[s][k] = role ? Qt.a[3 * k + s] : Pf.a[3 * k + s];
            t2[s][k] = role ? (S1.a[3 * k + s] + S3.a[3 * k + s]) + S2.a[3 * k + s] : -Qt.a[3 * s + k];
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        t1[3][k] = role ? 0.0 : RC(K_SO + k) + CY(6, k);
        t2[3][k] = role ? 0.0 : RC(K_SO + 3 + k) + CY(6, 3 + k);
    }
    double detA, detS, fA, fS;
    const M3 Ai = invDet(A, detA, fA);
    const M3 AiB = mul(Ai, Bm);
    const M3 S = subMulM(E, Cm, AiB);
    const M3 Si = invDet(S, detS, fS);
    double x1[4][3], x2[4][3];
    if (smallDet(detA, fA) || smallDet(detS, fS)) { // identical in both lanes of the beam (same D')
        double D36[36], cols[24];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                D36[6 * r + c] = A.a[3 * r + c]; D36[6 * r + c + 3] = Bm.a[3 * r + c];
                D36[6 * (r + 3) + c] = Cm.a[3 * r + c]; D36[6 * (r + 3) + c + 3] = E.a[3 * r + c];
            }
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) { cols[6 * s + k] = t1[s][k]; cols[6 * s + 3 + k] = t2[s][k]; }
        wsSolvePivoted(D36, cols);
#pragma unroll
        for (int s = 0; s < 4; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) { x1[s][k] = cols[6 * s + k]; x2[s][k] = cols[6 * s + 3 + k]; }
    } else {
#pragma unroll
        for (int s = 0; s < 4; s++) {
            const V3 y = mul(Ai, v3(t1[s][0], t1[s][1], t1[s][2]));
            const V3 b2 = mul(Si, subMul(v3(t2[s][0], t2[s][1], t2[s][2]), Cm, y));
            const V3 b1 = subMul(y, AiB, b2);
            x1[s][0] = b1.x; x1[s][1] = b1.y; x1[s][2] = b1.z;
            x2[s][0] = b2.x; x2[s][1] = b2.y; x2[s][2] = b2.z;
        }
    }
    // ---- uPrime columns / bPrime to the scratch (same layout as the one-thread-per-beam sweeps)
    if (role == 0) {
        const size_t b0 = tileRow(T, p.nmax, i, 6, lb);
#pragma unroll
        for (int k = 0; k < 3; k++) { stS(p.bP + b0 + (size_t)k * TS, x1[3][k]); stS(p.bP + b0 + (size_t)(3 + k) * TS, x2[3][k]); }
    }
    if (last) return;
    {
        double* u = p.uP + tileRow(T, p.nmax, i, 36, lb) + (size_t)(3 * role) * TS;
#pragma unroll
        for (int s = 0; s < 3; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                stS(u + (size_t)(6 * k + s) * TS, x1[s][k]);
                stS(u + (size_t)(6 * (k + 3) + s) * TS, x2[s][k]);
            }
    }
    // ---- the lane's columns of the carry of cell i+1:  N - l X,
    //      l = [[P, -Qt], [Qt.T(), Sl]], Sl = S1 + S3 - S2;  N = [[-P, -Qt], [-Qt.T(), S3 - S2 - S1]] | srcNei
    M3 Sl;
#pragma unroll
    for (int k = 0; k < 9; k++) Sl.a[k] = S1.a[k] + (S3.a[k] - S2.a[k]);
#pragma unroll
    for (int s = 0; s < 4; s++) {
        double n1[3], n2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (s < 3) {
                n1[k] = -t1[s][k];                                                                   // -P[:, s] or -Qt[:, s]
                n2[k] = role ? (S3.a[3 * k + s] - S2.a[3 * k + s]) - S1.a[3 * k + s] : t2[s][k];     // Sd[:, s] or -Qt[s, :]
            } else {
                n1[k] = RC(K_SN + k);
                n2[k] = RC(K_SN + 3 + k);
            }
        }
#pragma unroll
        for (int r = 0; r < 3; r++) {
            double ta = n1[r], ba = n2[r];
#pragma unroll
            for (int k = 0; k < 3; k++) ta = fma(-Pf.a[3 * r + k], x1[s][k], ta);
#pragma unroll
            for (int k = 0; k < 3; k++) ta = fma(Qt.a[3 * r + k], x2[s][k], ta);
#pragma unroll
            for (int k = 0; k < 3; k++) ba = fma(-Sl.a[3 * r + k], x2[s][k], ba);
#pragma unroll
            for (int k = 0; k < 3; k++) ba = fma(-Qt.a[3 * k + r], x1[s][k], ba);
            nc[s][r] = ta; nc[s][3 + r] = ba;
        }
    }
#undef RC
#undef CY
}

template <int SCHEME>
__global__ void __launch_bounds__(WS_THREADS) beam_forward_ws(const BeamParams p)
{
    extern __shared__ __align__(16) unsigned char ws_smem[];
    if (convLeave(p.cv, false)) return;
    double* sm = (double*)ws_smem;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane & 15, j = lane >> 4; // beam of the block; cell slot (assembler) / column role (eliminator)
    const int b = blockIdx.x * WS_BEAMS + g;
    const bool live = b < p.B;
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31;
    double* cy = sm + WS_CARRY_OFF + g;
    const int n = live ? p.nSeg[b] : 0;
    const int nW = __reduce_max_sync(0xffffffffu, n);
    const int nBatches = (nW + WS_BATCH - 1) / WS_BATCH;
    double res = 0.0;
    if (w == 1) {
        // ================= assembler warp: lane (j, g) assembles cell 2k + j of beam g into buffer k & 1
        const CoopBeam kb = coopLoadBeam<SCHEME>(p, b, live);
        wsPrefetchCell<SCHEME>(p, T, lb, WS_BATCH + j, n);
        for (int k = 0; k < nBatches; k++) {
            const int buf = k & 1, i = WS_BATCH * k + j;
            if (k >= 2) namedSync(BAR_EMPTY + buf); // the eliminator has finished with batch k - 2
            double* rec = sm + buf * 32 + lane;
            double so[6];
            const bool mine = i < n;
            if (mine) wsAssembleCell<SCHEME>(p, b, i, n, rec, cy, kb, so);
            wsPrefetchCell<SCHEME>(p, T, lb, i + 2 * WS_BATCH, n);
            __syncwarp();
            if (mine) { // ||source||^2 (EigenOF.C:253-282, x0 = 0): own part + neighbour part from face i (cell i-1's record)
                const double* prev = j ? rec - 16 : sm + (1 - buf) * 32 + 16 + g;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    const double s = so[r] + (i > 0 ? prev[(K_SN + r) * WS_COLS] : 0.0);
                    res = fma(s, s, res);
                }
            }
            __threadfence_block();
            namedArrive(BAR_FULL + buf);
        }
    } else {
        // ================= eliminator warp: lanes (role j, g) walk the recurrence of beam g
        for (int k = 0; k < nBatches; k++) {
            const int buf = k & 1;
            namedSync(BAR_FULL + buf);
#pragma unroll 1
            for (int jj = 0; jj < WS_BATCH; jj++) {
                const int ic = WS_BATCH * k + jj;
                const bool active = ic < n, last = ic == n - 1;
                double nc[4][6];
                if (active) wsEliminateCell(p, T, lb, ic, last, j, sm + buf * 32 + jj * 16 + g, cy, nc);
                __syncwarp(); // both lanes of every beam have read the old carry
                if (active && !last) {
#pragma unroll
                    for (int s = 0; s < 3; s++)
#pragma unroll
                        for (int r = 0; r < 6; r++) cy[(6 * (3 * j + s) + r) * WS_BEAMS] = nc[s][r];
                    if (j == 0) {
#pragma unroll
                        for (int r = 0; r < 6; r++) cy[(36 + r) * WS_BEAMS] = nc[3][r];
                    }
                }
                __syncwarp();
            }
            if (k + 2 < nBatches) namedArrive(BAR_EMPTY + buf);
        }
    }
    double v[1] = {res};
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, serial part: x_i = bPrime_i - uPrime_i & x_{i+1}.  One thread per beam; the 42 scratch rows of a cell are one
// contiguous 42 x 256 B run per warp tile, streamed through a ring of BS_RING cells by per-lane cp.async.
// X: tiled like a 6-component cell array, [tile][nmax][6][32].
// ---------------------------------------------------------------------------------------------------------------------
#define BS_THREADS 32
#define BS_RING 8
#define BS_SMEM_BYTES ((BS_THREADS / 32) * BS_RING * 42 * 32 * sizeof(double))
DEV void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
DEV void cpAsyncWaitGroup() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(BS_THREADS) beam_backsub(const BeamParams p, double* __restrict__ X)
{
    extern __shared__ __align__(16) unsigned char bs_smem[];
    if (convLeave(p.cv, true)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * BS_THREADS + threadIdx.x;
    const size_t T = (size_t)(b >> 5);
    const int Rc = p.nmax;
    const int n = b < p.B ? p.nSeg[b] : 0;
    const int nW = __reduce_max_sync(0xffffffffu, n);
    double* ring = (double*)bs_smem + (size_t)w * BS_RING * 42 * 32 + lane;
    auto issue = [&](int i) { // rows of cell i -> ring slot i % BS_RING (own column only: no barrier needed)
        if (i >= 0 && i < n) {
            double* dst = ring + (size_t)(i % BS_RING) * 42 * 32;
            const double* u = p.uP + tileRow(T, Rc, i, 36, lane);
            const double* bb = p.bP + tileRow(T, Rc, i, 6, lane);
            if (i < n - 1) {
#pragma unroll
                for (int k = 0; k < 36; k++) cpAsync8(dst + k * 32, u + (size_t)k * TS);
            }
#pragma unroll
            for (int k = 0; k < 6; k++) cpAsync8(dst + (36 + k) * 32, bb + (size_t)k * TS);
        }
        cpAsyncCommit(); // one group per cell index, also when this lane has nothing to copy
    };
    for (int d = 0; d < BS_RING - 1; d++) issue(nW - 1 - d);
    double xn[6] = {0, 0, 0, 0, 0, 0};
    for (int i = nW - 1; i >= 0; i--) {
        issue(i - (BS_RING - 1));
        cpAsyncWaitGroup<BS_RING - 1>(); // the group of cell i has landed
        if (i < n) {
            const double* src = ring + (size_t)(i % BS_RING) * 42 * 32;
            double x[6];
#pragma unroll
            for (int r = 0; r < 6; r++) x[r] = src[(36 + r) * 32];
            if (i < n - 1) {
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) acc += src[(6 * r + c) * 32] * xn[c];
                    x[r] -= acc;
                }
            }
            double* o = X + tileRow(T, Rc, i, 6, lane);
#pragma unroll
            for (int r = 0; r < 6; r++) { xn[r] = x[r]; o[(size_t)r * TS] = x[r]; }
        }
    }
}

// The same after a SPLIT forward sweep (beam_forward_split: the left half eliminated left to right, the mirrored half right
// to left): two threads per beam, grid (ceil(B / BS_THREADS), 2).  Both solve the 6x6 interface system (splitInterface), then
// the left half substitutes outwards to cell 0, x_i = bPrime_i - uPrime_i x_{i+1}, and the mirrored half to cell n-1,
// x_i = bPrime_i - uPrime_i x_{i-1}.  Uniform bundles (every beam nmax cells).
__global__ void __launch_bounds__(BS_THREADS) beam_backsub_split(const BeamParams p, double* __restrict__ X)
{
    extern __shared__ __align__(16) unsigned char bs_smem[];
    if (convLeave(p.cv, true)) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int b = blockIdx.x * BS_THREADS + threadIdx.x;
    const bool mir = blockIdx.y != 0;
    const size_t T = (size_t)(b >> 5);
    const int Rc = p.nmax, n = p.nmax, hL = (n + 1) / 2;
    const bool live = b < p.B;
    const int m = mir ? n - hL : hL; // steps of this half; step k visits cell hL + k (mirrored) or hL - 1 - k
    double* ring = (double*)bs_smem + (size_t)w * BS_RING * 42 * 32 + lane;
    auto cellOf = [&](int k) { return mir ? hL + k : hL - 1 - k; };
    auto issue = [&](int k) { // rows of step k -> ring slot k % BS_RING (own column only: no barrier needed)
        if (live && k >= 1 && k < m) { // step 0 takes its x from the interface solve
            const int i = cellOf(k);
            double* dst = ring + (size_t)(k % BS_RING) * 42 * 32;
            const double* u = p.uP + tileRow(T, Rc, i, 36, lane);
            const double* bb = p.bP + tileRow(T, Rc, i, 6, lane);
#pragma unroll
            for (int r = 0; r < 36; r++) cpAsync8(dst + r * 32, u + (size_t)r * TS);
#pragma unroll
            for (int r = 0; r < 6; r++) cpAsync8(dst + (36 + r) * 32, bb + (size_t)r * TS);
        }
        cpAsyncCommit();
    };
    for (int d = 0; d < BS_RING - 1; d++) issue(d);
    double xL[6], xR[6], xn[6];
    if (live) splitInterface(p, T, lane, hL, xL, xR);
    for (int k = 0; k < m; k++) {
        issue(k + BS_RING - 1);
        cpAsyncWaitGroup<BS_RING - 1>();
        if (!live) continue;
        double x[6];
        if (k == 0) {
#pragma unroll
            for (int r = 0; r < 6; r++) x[r] = mir ? xR[r] : xL[r];
        } else {
            const double* src = ring + (size_t)(k % BS_RING) * 42 * 32;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 6; c++) acc += src[(6 * r + c) * 32] * xn[c];
                x[r] = src[(36 + r) * 32] - acc;
            }
        }
        double* o = X + tileRow(T, Rc, cellOf(k), 6, lane);
#pragma unroll
        for (int r = 0; r < 6; r++) { xn[r] = x[r]; o[(size_t)r * TS] = x[r]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, parallel part: one thread per (tile, cell, lane) = per cell of the bundle, beams fastest (coalesced rows).
// ---------------------------------------------------------------------------------------------------------------------
DEV bool cellOfThread(const BeamParams& p, int& b, int& i)
{
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(t & 31);
    const size_t row = t >> 5; // T * nmax + i
    const size_t T = row / (size_t)p.nmax;
    i = (int)(row - T * (size_t)p.nmax);
    b = (int)(T * 32) + lane;
    return b < p.B && i < p.nSeg[b];
}
DEV void ldX(const double* __restrict__ X, size_t T, int Rc, int i, int lane, V3& a, V3& c)
{
    const size_t x0 = tileRow(T, Rc, i, 6, lane);
    a = v3(X[x0], X[x0 + TS], X[x0 + 2 * TS]);
    c = v3(X[x0 + 3 * TS], X[x0 + 4 * TS], X[x0 + 5 * TS]);
}
// Faces first: they read the neighbour cell's pre-update W ...
#ifndef BEAM_FACES_MIN_BLOCKS
#define BEAM_FACES_MIN_BLOCKS 3 // 168 registers: 12 warps per SM; measured 6 % faster than 2 (255 registers) or 4 (128, spills) at 4096 beams
#endif
__global__ void __launch_bounds__(128, BEAM_FACES_MIN_BLOCKS) beam_update_faces(const BeamParams p, const double* __restrict__ X)
{
    if (convLeave(p.cv, true)) return;
    int b, i;
    if (!cellOfThread(p, b, i)) return;
    const size_t T = (size_t)(b >> 5), sb = p.Bpad;
    const int lane = b & 31, Rc = p.nmax, n = p.nSeg[b];
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ, delta = 1.0 / dcB;
    const V3 CQ = ld3(p.CQ, b, sb), CM = ld3(p.CM, b, sb);
    const M3 refL = ld9(p.refL, b, sb);
    const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]), e0 = mulT(refL, dR0);
    V3 DW, DT, xWn = v3(0, 0, 0), xTn = v3(0, 0, 0), WoldN = v3(0, 0, 0);
    ldX(X, T, Rc, i, lane, DW, DT);
    if (i < n - 1) {
        ldX(X, T, Rc, i + 1, lane, xWn, xTn);
        WoldN = ld3(p.W, tileRow(T, Rc, i + 1, 3, lane), TS);
    }
    updateFacesOfCell(p, b, i, n, CQ, CM, dR0, e0, dcI, dcB, delta, ld3(p.W, tileRow(T, Rc, i, 3, lane), TS), WoldN, DW, DT, xWn, xTn);
}
// ... then the cells, and the two solution norms.
template <int SCHEME>
__global__ void __launch_bounds__(128) beam_update_cells(const BeamParams p, const double* __restrict__ X)
{
    if (convLeave(p.cv, true)) return;
    int b, i;
    double v[2] = {0.0, 0.0};
    if (cellOfThread(p, b, i)) {
        V3 DW, DT, Wold;
        ldX(X, (size_t)(b >> 5), p.nmax, i, b & 31, DW, DT);
        v[0] = DW.x * DW.x + DW.y * DW.y + DW.z * DW.z + DT.x * DT.x + DT.y * DT.y + DT.z * DT.z;
        updateCell<SCHEME>(p, (size_t)(b >> 5), b & 31, i, DW, DT, Wold, v[1]);
    }
    blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.x);
}
