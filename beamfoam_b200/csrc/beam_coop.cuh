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
//   beam_update      updateSolutionVariables (Evolve.C:757-923) and the patch evaluate()s with one thread per CELL over the
//                    whole bundle: no serial dependence once x is known, so the latency of this (HBM-bound) part is hidden by
//                    occupancy, not by the length of a beam.  A thread updates its cell, the face towards the next cell (with
//                    the neighbour's pre-update W, which the forward part saved) and the end patches of its beam; norms.
//
// State, scratch layout (uPrime/bPrime) and every per-cell / per-face formula are shared with the one-thread-per-beam
// kernels, so the two forward and the two backward implementations are interchangeable (tests/test_gpu_coop.py runs all
// four combinations against the oracle).
#pragma once
#include "beam_kernels.cuh"

#define WS_THREADS(EW) (32 * ((EW) + 1))
// Template parameters of the forward kernel:
//   BPB   beams (or half beams) per block, 16 or 32.  16: the assembler warp takes 2 consecutive cells x 16 beams per pass and
//         the eliminator warp gives every beam 2 lanes (columns 0..2 + right-hand side | columns 3..5): the lowest latency per
//         cell, twice the warps.  32: one lane per beam in both warps (the assembler works one cell ahead of the eliminator, which
//         solves all 7 columns): half the warps per beam -- the family for bundles that would not fit the GPU with 16.
//   SPLIT the block serves one HALF of each of its beams (blockIdx.y = 0 left halves, 1 mirrored halves), as in
//         beam_forward_split: the two halves eliminate towards the middle face, the interface is solved by the backward part.
//   EW    eliminator warps.  2 (with BPB = 32): the two column groups of a beam are two WARPS (warp 0: columns 0..2 and the
//         right-hand side, warp 1: columns 3..5; warp 2 assembles) instead of two lanes of one warp -- the lane efficiency of
//         BPB = 32 with the per-cell latency of BPB = 16.  The carry is double-buffered and the two eliminator warps meet at
//         one named barrier per cell.
#define WS_BATCH(BPB) (32 / (BPB)) // cells per beam per assembler pass
// record rows (per cell; row k of column col at rec[k * WS_COLS + col], col = buffer * 32 + slot * BPB + beam)
enum {
    K_DI = 0,   // 1: QRhoCoeff (inertia part of the translational diagonal, Evolve.C:459-535)
    K_EO = 1,   // 9: own part of the rotational diagonal block: ((S3 - S1) + S2) + MRhoCoeff; S2 = -spin(hM)
    K_SO = 10,  // 6: own part of source_i (face i+1, loads, inertia, end patches)
    K_SN = 16,  // 6: neighbour part of source_{i+1} from face i+1: (eQ ; eM - eMQ)
    // the blocks of u_i = [[P, Qt], [-Qt.T(), Su]] and of the neighbour part N = [[-P, -Qt], [-Qt.T(), Sd]] of d_{i+1}, laid out
    // so that the column group of a lane (role 0: columns 0..2, role 1: columns 3..5) is an ADDRESS offset, not a select:
    K_P = 22,   // 9: P = a CQW                                  u11          (+ 9 role -> u12)
    K_QT = 31,  // 9: Qt = 0.5 CQTheta                           u12
    K_NQT = 40, // 9: -Qt.T()                                    u21 = N21    (+ 9 role -> u22, + 18 role -> N22)
    K_SU = 49,  // 9: Su = (S1 + S3) + S2                        u22          S1 = a CMTheta, S3 = 0.5 CMQTheta
    K_SD = 58,  // 9: Sd = (S3 - S2) - S1                        N22
    K_SL = 67,  // 9: Sl = S1 + (S3 - S2)                        l22
    K_ROWS = 76,
    K_PD = K_SN // the LAST cell of a beam has no face i+1: rows K_SN .. K_SN+35 hold the right-patch closure (6x6, row-major)
};
#define WS_COLS 64
#define WS_CARRY_OFF (K_ROWS * WS_COLS)        // carry [7 columns][6 rows][BPB beams]
#define WS_SMEM_BYTES(BPB, EW) ((WS_CARRY_OFF + 42 * (BPB) * (EW)) * sizeof(double))
enum { BAR_FULL = 1, BAR_EMPTY = 3, BAR_CARRY = 5 }; // named barriers 1,2 / 3,4 / 5 (0 is __syncthreads)

DEV void namedSync(int id, int nThreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nThreads) : "memory"); }
DEV void namedArrive(int id, int nThreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nThreads) : "memory"); }

// Pivoted fallback of the elimination: D (row-major 36) x = cols, NS right-hand sides per lane.
template <int NS>
__device__ __noinline__ void wsSolvePivoted(const double* __restrict__ D36, double* __restrict__ cols)
{
    double A[6][6], Bq[6][NS];
#pragma unroll
    for (int r = 0; r < 6; r++) {
#pragma unroll
        for (int c = 0; c < 6; c++) A[r][c] = D36[6 * r + c];
#pragma unroll
        for (int s = 0; s < NS; s++) Bq[r][s] = cols[6 * s + r];
    }
    solve6<NS>(A, Bq);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int s = 0; s < NS; s++) cols[6 * s + r] = Bq[r][s];
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

// L2 prefetch of what the assembly of the cell in row `row` of beam (T, lb) reads; faceRow = the face towards the next cell of
// the walk (row + 1, or row for a mirrored half), negative when the cell has none
template <int SCHEME>
DEV void wsPrefetchCell(const BeamParams& p, const size_t T, const int lb, const int row, const int faceRow)
{
    const int Rc = p.nmax, Rf = p.nmax + 1;
    const size_t c3 = tileRow(T, Rc, row, 3, lb), c4 = tileRow(T, Rc, row, 4, lb);
    pfRows<3>(p.W, c3, TS);
    if (faceRow >= 0) {
        const size_t f3 = tileRow(T, Rf, faceRow, 3, lb);
        pfRows<4>(p.Lf, tileRow(T, Rf, faceRow, 4, lb), TS);
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
// row: the cell's row in the tile.  firstCell: the first cell of the walk -- the carry of the beam (cyInit, stride BPB) is
// initialised with the closure of the end patch the walk starts from (the left patch; the right patch for a mirrored half).
// hasFace: the cell has an internal face towards the next cell of the walk (row + 1; row - 1 for a mirrored half, whose face
// is seen from its neighbour side: the same code with th, hQ, hM, eQ, eM negated, see fwdFacePart<.., MIR>); otherwise it is
// the last cell of an unsplit beam and closes the right patch.  Returns the own part of the source (K_SO) in so[]; the
// complete source of the cell is so + SN of the previous cell of the walk.
template <int SCHEME, bool MIR, int BPB>
DEV void wsAssembleCell(const BeamParams& p, const int b, const int row, const bool firstCell, const bool hasFace,
                        double* __restrict__ rec, double* __restrict__ cyInit, const CoopBeam& kb, double (&so)[6])
{
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31, Rc = p.nmax, Rf = p.nmax + 1;
    const CellIn cell = loadCellIn<SCHEME>(p, b, row); // issued first: consumed after the face part
    double src[6];
#pragma unroll
    for (int r = 0; r < 6; r++) src[r] = 0.0;
    const V3 Wi = ld3(p.W, tileRow(T, Rc, row, 3, lb), TS);
    if (p.wSave) st3(p.wSave, tileRow(T, Rc, row, 3, lb), TS, Wi); // this cell's pre-update W, for beam_update
    if (firstCell) { // end patch (BC.C:64-1190): its 6x6 part starts the carry, its source part joins the own source
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        patchClosure(p, MIR ? 1 : 0, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            src[r] = s6[r];
            cyInit[(36 + r) * BPB] = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) cyInit[(6 * c + r) * BPB] = D36[6 * r + c]; // the carry is column-major
        }
    }
    double eo[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; // own part of the rotational diagonal block from the face (none for a last cell)
    if (hasFace) { // internal face towards the next cell: the same folded form as fwdFacePart (beam_kernels.cuh)
        const int rowN = MIR ? row - 1 : row + 1, fr = MIR ? row : row + 1;
        const size_t f3 = tileRow(T, Rf, fr, 3, lb);
        const V3 Wn = ld3(p.W, tileRow(T, Rc, rowN, 3, lb), TS);
        const Q4 qf = ldq(p.Lf, tileRow(T, Rf, fr, 4, lb), TS);
        const V3 Kf = ld3(p.K, f3, TS), Qf = ld3(p.Q, f3, TS), Mf = ld3(p.M, f3, TS);
        const double h = 0.5 * kb.dZ, dcI = kb.dcI, sdc = MIR ? -dcI : dcI;
        const V3 t = v3(kb.dR0.x + sdc * (Wn.x - Wi.x), kb.dR0.y + sdc * (Wn.y - Wi.y), kb.dR0.z + sdc * (Wn.z - Wi.z));
        if (MIR && p.ifW && row == p.hL) st3(p.ifW, b, p.Bpad, Wn); // pre-update W of cell h-1 (beam_backward_split)
        const V3 th = (MIR ? -h : h) * t;
        const M3 Lm = quatToMat(qf);
        const V3 hQ = (MIR ? -0.5 : 0.5) * Qf;
        const Sym3 Pm = conjDiagSym(Lm, dcI * kb.CQ);                           // a*CQW :680
        V3 eQ = mul(Lm, cmul(kb.CQ, mulTSub(Lm, t, v3(1.0, 0.0, 0.0))));        // :682 with Gamma of :881-887
        if (MIR) eQ = -eQ;
        const M3 Gh = mulSpinR(Pm, th);
        M3 Qt = Gh;                                                             // 0.5*CQTheta :686-693
        Qt.a[1] += hQ.z; Qt.a[2] -= hQ.y; Qt.a[3] -= hQ.z; Qt.a[5] += hQ.x; Qt.a[6] += hQ.y; Qt.a[7] -= hQ.x;
        const V3 eMQ = cross(th, eQ);                                           // :737-741
        const Sym3 S1 = conjDiagSym(Lm, dcI * kb.CM);                           // a*CMTheta :699
        V3 eM = mul(Lm, cmul(kb.CM, Kf));                                       // :684
        if (MIR) eM = -eM;
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
        const V3 hM = (MIR ? -0.5 : 0.5) * Mf;                                  // 0.5*CMTheta2 = -S(hM) :703
        const M3 Pf = full(Pm), S1f = full(S1);
        subV(src, 0, eQ);
        subV(src, 3, eM + eMQ);
        const V3 sn2 = eM - eMQ;
        rec[(K_SN + 0) * WS_COLS] = eQ.x; rec[(K_SN + 1) * WS_COLS] = eQ.y; rec[(K_SN + 2) * WS_COLS] = eQ.z;
        rec[(K_SN + 3) * WS_COLS] = sn2.x; rec[(K_SN + 4) * WS_COLS] = sn2.y; rec[(K_SN + 5) * WS_COLS] = sn2.z;
        const double s2[9] = {0.0, hM.z, -hM.y, -hM.z, 0.0, hM.x, hM.y, -hM.x, 0.0}; // S2 = -spin(hM)
#pragma unroll
        for (int k = 0; k < 9; k++) {
            const double sm3 = S3.a[k] - s2[k]; // S3 - S2
            rec[(K_P + k) * WS_COLS] = Pf.a[k];
            rec[(K_QT + k) * WS_COLS] = Qt.a[k];
            rec[(K_NQT + k) * WS_COLS] = -Qt.a[3 * (k % 3) + k / 3];
            rec[(K_SU + k) * WS_COLS] = (S1f.a[k] + S3.a[k]) + s2[k];
            rec[(K_SD + k) * WS_COLS] = sm3 - S1f.a[k];
            rec[(K_SL + k) * WS_COLS] = S1f.a[k] + sm3;
            eo[k] = (S3.a[k] - S1f.a[k]) + s2[k];
        }
    } else { // the last cell of an unsplit beam: right patch instead of a face; its 6x6 part takes the tensor rows of the record
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
    addLoadsAndInertia<SCHEME>(p, b, row, kb.dZ, kb.ARho, kb.CI, cell, D, src); // touches D[0..2][0..2] (diagonal) and D[3..5][3..5]
    rec[K_DI * WS_COLS] = D[0][0];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) rec[(K_EO + 3 * r + c) * WS_COLS] = hasFace ? eo[3 * r + c] + D[3 + r][3 + c] : D[3 + r][3 + c];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        rec[(K_SO + r) * WS_COLS] = src[r];
        so[r] = src[r];
    }
}

// ---- eliminator: one cell by the ROLES lanes of its beam -----------------------------------------------------------------
// rc = record column of the cell, cy = carry column of the beam (+ g, stride BPB), row = the cell's row in the tile (where
// uPrime / bPrime go).  ROLES = 2: role 0 solves for columns 0..2 of u_i and the right-hand side, role 1 for columns 3..5 (and
// an idle fourth slot).  ROLES = 1: the lane solves for all six columns and the right-hand side.  Returns the lane's columns of
// the new carry in nc[slot][row].
template <int ROLES, int BPB>
DEV void wsEliminateCell(const BeamParams& p, const size_t T, const int lb, const int row, const bool last, const int role,
                         const double* __restrict__ rc, const double* __restrict__ cy, double (&nc)[ROLES == 2 ? 4 : 7][6])
{
    constexpr int NU = ROLES == 2 ? 3 : 6, NS = NU + 1; // u columns of the lane; slot NU = the right-hand side
#define RC(k) rc[(k) * WS_COLS]
#define CY(c, r) cy[(6 * (c) + (r)) * BPB]
#define UPPER(s) (ROLES == 2 ? role : (s) >= 3) // the slot's column is one of columns 3..5 (ROLES = 2: an address offset, see the record)
    M3 Pf, Qt;
    M3 A, Bm, Cm, E;
    const double qd = RC(K_DI);
    if (!last) {
#pragma unroll
        for (int k = 0; k < 9; k++) { Pf.a[k] = RC(K_P + k); Qt.a[k] = RC(K_QT + k); }
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { // D'_i = carry (column-major) + own part: [[-P + q I, Qt], [Qt.T(), S3 - S1 + S2 + inertia]]
                A.a[3 * r + c] = CY(c, r) - Pf.a[3 * r + c] + (r == c ? qd : 0.0);
                Bm.a[3 * r + c] = CY(c + 3, r) + Qt.a[3 * r + c];
                Cm.a[3 * r + c] = CY(c, r + 3) + Qt.a[3 * c + r];
                E.a[3 * r + c] = CY(c + 3, r + 3) + RC(K_EO + 3 * r + c);
            }
    } else {
        Pf = Qt = m3zero();
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) { // carry + right-patch closure + inertia
                A.a[3 * r + c] = CY(c, r) + RC(K_PD + 6 * r + c) + (r == c ? qd : 0.0);
                Bm.a[3 * r + c] = CY(c + 3, r) + RC(K_PD + 6 * r + c + 3);
                Cm.a[3 * r + c] = CY(c, r + 3) + RC(K_PD + 6 * (r + 3) + c);
                E.a[3 * r + c] = CY(c + 3, r + 3) + (RC(K_PD + 6 * (r + 3) + c + 3) + RC(K_EO + 3 * r + c));
            }
    }
    // ---- the lane's right-hand sides: slots 0..NU-1 = its columns of u, slot NU = source - l bPrime (role 0 only)
    //      u = [[P, Qt], [-Qt.T(), Su]]: column q of the lane's group sits 9 * (group) rows further down the record
    double t1[NS][3], t2[NS][3];
#pragma unroll
    for (int s = 0; s < NU; s++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const int q = s % 3, up = UPPER(s) ? 9 : 0;
            t1[s][k] = RC(K_P + up + 3 * k + q); // (a last cell reads rows of its patch closure here: its u columns are solved
            t2[s][k] = RC(K_NQT + up + 3 * k + q); //  for but never stored)
        }
#pragma unroll
    for (int k = 0; k < 3; k++) {
        t1[NU][k] = role ? 0.0 : RC(K_SO + k) + CY(6, k);
        t2[NU][k] = role ? 0.0 : RC(K_SO + 3 + k) + CY(6, 3 + k);
    }
    double detA, detS, fA, fS;
    const M3 Ai = invDet(A, detA, fA);
    const M3 AiB = mul(Ai, Bm);
    const M3 S = subMulM(E, Cm, AiB);
    const M3 Si = invDet(S, detS, fS);
    double x1[NS][3], x2[NS][3];
    if (smallDet(detA, fA) || smallDet(detS, fS)) { // identical in all lanes of the beam (same D')
        double D36[36], cols[6 * NS];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                D36[6 * r + c] = A.a[3 * r + c]; D36[6 * r + c + 3] = Bm.a[3 * r + c];
                D36[6 * (r + 3) + c] = Cm.a[3 * r + c]; D36[6 * (r + 3) + c + 3] = E.a[3 * r + c];
            }
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) { cols[6 * s + k] = t1[s][k]; cols[6 * s + 3 + k] = t2[s][k]; }
        wsSolvePivoted<NS>(D36, cols);
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) { x1[s][k] = cols[6 * s + k]; x2[s][k] = cols[6 * s + 3 + k]; }
    } else {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const V3 y = mul(Ai, v3(t1[s][0], t1[s][1], t1[s][2]));
            const V3 b2 = mul(Si, subMul(v3(t2[s][0], t2[s][1], t2[s][2]), Cm, y));
            const V3 b1 = subMul(y, AiB, b2);
            x1[s][0] = b1.x; x1[s][1] = b1.y; x1[s][2] = b1.z;
            x2[s][0] = b2.x; x2[s][1] = b2.y; x2[s][2] = b2.z;
        }
    }
    // ---- uPrime columns / bPrime to the scratch (same layout as the one-thread-per-beam sweeps)
    if (role == 0) {
        const size_t b0 = tileRow(T, p.nmax, row, 6, lb);
#pragma unroll
        for (int k = 0; k < 3; k++) { stS(p.bP + b0 + (size_t)k * TS, x1[NU][k]); stS(p.bP + b0 + (size_t)(3 + k) * TS, x2[NU][k]); }
    }
    if (last) return;
    {
        double* u = p.uP + tileRow(T, p.nmax, row, 36, lb) + (size_t)(3 * role) * TS;
#pragma unroll
        for (int s = 0; s < NU; s++)
#pragma unroll
            for (int k = 0; k < 3; k++) {
                stS(u + (size_t)(6 * k + s) * TS, x1[s][k]);
                stS(u + (size_t)(6 * (k + 3) + s) * TS, x2[s][k]);
            }
    }
    // ---- the lane's columns of the carry of the next cell:  N - l X,
    //      l = [[P, -Qt], [Qt.T(), Sl]], Sl = S1 + S3 - S2;  N = [[-P, -Qt], [-Qt.T(), S3 - S2 - S1]] | srcNei
    M3 Sl;
#pragma unroll
    for (int k = 0; k < 9; k++) Sl.a[k] = RC(K_SL + k);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        double n1[3], n2[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
            if (s < NU) {
                const int q = s % 3, up = UPPER(s) ? 18 : 0;
                n1[k] = -t1[s][k];                      // -P[:, q] or -Qt[:, q]
                n2[k] = RC(K_NQT + up + 3 * k + q);     // -Qt[q, :] or Sd[:, q]
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
#undef UPPER
}

// grid (ceil(Bpad / BPB), SPLIT ? 2 : 1), WS_THREADS(EW) threads
template <int SCHEME, int BPB, bool SPLIT, int EW>
__global__ void __launch_bounds__(WS_THREADS(EW), EW == 2 ? 4 : 1) beam_forward_ws(const __grid_constant__ BeamParams p)
{
    static_assert(EW == 1 || BPB == 32, "two eliminator warps serve one tile of 32 beams");
    constexpr int NB = WS_BATCH(BPB), ROLES = EW == 2 ? 2 : 32 / BPB, NU = ROLES == 2 ? 3 : 6, NT = WS_THREADS(EW);
    extern __shared__ __align__(16) unsigned char ws_smem[];
    if (convLeave(p.cv, false)) return;
    double* sm = (double*)ws_smem;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int g = lane % BPB, j = lane / BPB; // beam of the block; cell slot (assembler) / column role (one eliminator warp)
    const int role = EW == 2 ? w : j;
    const int b = blockIdx.x * BPB + g;
    const bool live = b < p.B, mir = SPLIT && blockIdx.y != 0;
    const size_t T = (size_t)(b >> 5);
    const int lb = b & 31;
    double* cy = sm + WS_CARRY_OFF + g; // EW = 2: two buffers, cell ic reads buffer ic & 1 and writes the other
    const int n = live ? p.nSeg[b] : 0;
    // cells of this lane's walk: the whole beam, or its half (left: rows 0 .. hL-1 upwards, mirrored: rows nmax-1 .. hL downwards)
    const int hL = p.hL;
    const int m = !SPLIT ? n : !live ? 0 : mir ? p.nmax - hL : hL;
    const int nW = __reduce_max_sync(0xffffffffu, m);
    const int nBatches = (nW + NB - 1) / NB;
    auto rowOf = [&](int il) { return mir ? p.nmax - 1 - il : il; };
    auto faceOf = [&](int il) { return (SPLIT || il < n - 1) ? (mir ? rowOf(il) : il + 1) : -1; };
    double res = 0.0;
    if (w == EW) {
        // ================= assembler warp: lane (j, g) assembles local cell NB k + j of beam g into buffer k & 1
        const CoopBeam kb = coopLoadBeam<SCHEME>(p, b, live);
        if (NB + j < m) wsPrefetchCell<SCHEME>(p, T, lb, rowOf(NB + j), faceOf(NB + j));
        for (int k = 0; k < nBatches; k++) {
            const int buf = k & 1, il = NB * k + j;
            if (k >= 2) namedSync(BAR_EMPTY + buf, NT); // the eliminators have finished with batch k - 2
            double* rec = sm + buf * 32 + lane;
            double so[6];
            const bool mine = il < m;
            if (mine) {
                const bool hasFace = SPLIT || il < n - 1;
                if (mir) wsAssembleCell<SCHEME, true, BPB>(p, b, rowOf(il), il == 0, hasFace, rec, cy, kb, so);
                else wsAssembleCell<SCHEME, false, BPB>(p, b, rowOf(il), il == 0, hasFace, rec, cy, kb, so);
            }
            if (il + 2 * NB < m) wsPrefetchCell<SCHEME>(p, T, lb, rowOf(il + 2 * NB), faceOf(il + 2 * NB));
            __syncwarp();
            if (mine) { // ||source||^2 (EigenOF.C:253-282, x0 = 0): own part + neighbour part from the previous cell's record
                const double* prev = j ? rec - BPB : sm + (1 - buf) * 32 + (NB - 1) * BPB + g;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    const double s = so[r] + (il > 0 ? prev[(K_SN + r) * WS_COLS] : 0.0);
                    res = fma(s, s, res);
                }
            }
            __threadfence_block();
            namedArrive(BAR_FULL + buf, NT);
        }
    } else {
        // ================= eliminator warp(s): lanes (role, g) walk the recurrence of beam g
        for (int k = 0; k < nBatches; k++) {
            const int buf = k & 1;
            namedSync(BAR_FULL + buf, NT);
#pragma unroll 1
            for (int jj = 0; jj < NB; jj++) {
                const int ic = NB * k + jj;
                const bool active = ic < m, last = !SPLIT && ic == n - 1;
                const double* cyR = cy + (EW == 2 ? (ic & 1) * 42 * BPB : 0);
                double* cyW = cy + (EW == 2 ? ((ic + 1) & 1) * 42 * BPB : 0);
                double nc[NU + 1][6];
                if (active) wsEliminateCell<ROLES, BPB>(p, T, lb, rowOf(ic), last, role, sm + buf * 32 + jj * BPB + g, cyR, nc);
                if (EW == 1) __syncwarp(); // all lanes of every beam have read the old carry
                if (active && !last) {
#pragma unroll
                    for (int s = 0; s < NU; s++)
#pragma unroll
                        for (int r = 0; r < 6; r++) cyW[(6 * (3 * role + s) + r) * BPB] = nc[s][r];
                    if (role == 0) {
#pragma unroll
                        for (int r = 0; r < 6; r++) cyW[(36 + r) * BPB] = nc[NU][r];
                    }
                }
                if (EW == 1) __syncwarp();
                else namedSync(BAR_CARRY, 64); // both column groups of the new carry are in place
            }
            if (k + 2 < nBatches) namedArrive(BAR_EMPTY + buf, NT);
        }
    }
    double v[1] = {res};
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.y * gridDim.x + blockIdx.x);
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

__global__ void __launch_bounds__(BS_THREADS) beam_backsub(const __grid_constant__ BeamParams p, double* __restrict__ X)
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

// The same after a SPLIT forward sweep (beam_forward_split / beam_forward_ws<.., SPLIT>: the left half eliminated left to
// right, the mirrored half right to left): two threads per beam, grid (Bpad / 32, 2), one warp (= one tile of 32 beams) per
// block.  Both halves solve the 6x6 interface system (splitInterface), then the left half substitutes outwards to cell 0,
// x_i = bPrime_i - uPrime_i x_{i+1}, and the mirrored half to cell n-1, x_i = bPrime_i - uPrime_i x_{i-1}.  Uniform bundles.
// The 42 scratch rows of a cell are one contiguous run per array in the tile (36 x 256 B and 6 x 256 B): the TMA engine brings
// them in (two bulk copies per cell, issued by lane 0, completion on the mbarrier of the ring slot) while every lane pulls the
// lines of the cell BSS_PF steps ahead into L2.  A step is then 42 LDS + 36 FMA + 6 STG per lane -- the cp.async ring this
// replaces spent ~170 instructions per lane and step on issuing the copies (59 us for 100 steps at 4096 beams).
#define BSS_RING 4
#define BSS_PF 10
#define BSS_SLOT (42 * 32)
#define BSS_SMEM_BYTES (128 + BSS_RING * BSS_SLOT * sizeof(double))
__global__ void __launch_bounds__(32) beam_backsub_split(const __grid_constant__ BeamParams p, double* __restrict__ X)
{
    extern __shared__ __align__(128) unsigned char bs_smem[];
    if (convLeave(p.cv, true)) return;
    uint64_t* bars = (uint64_t*)bs_smem;
    double* ring = (double*)(bs_smem + 128);
    const int lane = threadIdx.x;
    const int b = blockIdx.x * 32 + lane;
    const bool mir = blockIdx.y != 0;
    const size_t T = blockIdx.x;
    const int Rc = p.nmax, n = p.nmax, hL = p.hL;
    const bool live = b < p.B;
    const int m = mir ? n - hL : hL; // steps of this half; step k visits cell hL + k (mirrored) or hL - 1 - k
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < BSS_RING; r++) mbarInit(bars + r, 1);
    }
    __syncwarp();
    auto cellOf = [&](int k) { return mir ? hL + k : hL - 1 - k; };
    auto issue = [&](int k) { // lane 0: rows of step k -> ring slot k % BSS_RING
        if (k >= 1 && k < m) { // step 0 takes its x from the interface solve
            const int i = cellOf(k), slot = k % BSS_RING;
            mbarExpectTx(bars + slot, 42 * 256);
            bulkLoad(ring + (size_t)slot * BSS_SLOT, p.uP + tileRow(T, Rc, i, 36, 0), 36 * 256, bars + slot);
            bulkLoad(ring + (size_t)slot * BSS_SLOT + 36 * 32, p.bP + tileRow(T, Rc, i, 6, 0), 6 * 256, bars + slot);
        }
    };
    auto prefetch = [&](int k) { // every lane: the 72 + 12 lines (128 B) of step k into L2
        if (k >= 1 && k < m) {
            const int i = cellOf(k);
            const double* u = p.uP + tileRow(T, Rc, i, 36, 0) + 16 * lane;
            pfL2(u); pfL2(u + 512);
            if (lane < 8) pfL2(u + 1024);
            if (lane < 12) pfL2(p.bP + tileRow(T, Rc, i, 6, 0) + 16 * lane);
        }
    };
    for (int k = 1; k <= BSS_PF; k++) prefetch(k);
    if (lane == 0)
        for (int k = 1; k <= BSS_RING; k++) issue(k);
    double xL[6], xR[6], xn[6];
    splitInterface(p, T, lane, hL, xL, xR);
#pragma unroll
    for (int r = 0; r < 6; r++) xn[r] = mir ? xR[r] : xL[r];
    if (live) {
        double* o = X + tileRow(T, Rc, cellOf(0), 6, lane);
#pragma unroll
        for (int r = 0; r < 6; r++) o[(size_t)r * TS] = xn[r];
    }
    for (int k = 1; k < m; k++) {
        const int slot = k % BSS_RING;
        prefetch(k + BSS_PF);
        mbarWait(bars + slot, (uint32_t)((k - 1) / BSS_RING) & 1u); // use j of a slot completes phase j: steps 1..R are use 0
        const double* src = ring + (size_t)slot * BSS_SLOT + lane;
        double u[36], x[6];
#pragma unroll
        for (int r = 0; r < 36; r++) u[r] = src[r * 32];
#pragma unroll
        for (int r = 0; r < 6; r++) x[r] = src[(36 + r) * 32];
        fenceProxyAsync();
        __syncwarp(); // every lane holds the slot in registers: it can be refilled
        if (lane == 0) issue(k + BSS_RING);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            double acc = 0.0;
#pragma unroll
            for (int c = 0; c < 6; c++) acc += u[6 * r + c] * xn[c];
            x[r] -= acc;
        }
        if (live) {
            double* o = X + tileRow(T, Rc, cellOf(k), 6, lane);
#pragma unroll
            for (int r = 0; r < 6; r++) o[(size_t)r * TS] = x[r];
        }
#pragma unroll
        for (int r = 0; r < 6; r++) xn[r] = x[r];
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, parallel part: one thread per (tile, cell, lane) = per cell of the bundle, beams fastest (coalesced rows).
// ---------------------------------------------------------------------------------------------------------------------
// updateSolutionVariables (Evolve.C:757-923) and the patch evaluate()s for ONE cell and the faces it owns, one thread per cell.
// The face between cells i and i+1 needs the pre-update W of BOTH cells while the thread of cell i+1 is writing its new W: the
// forward part of the iteration saved every cell's old W (BeamParams::wSave, three streaming stores per cell from a register it
// holds anyway), so the two updates have no read-after-write hazard and run in one kernel -- one launch and 9 of 73 doubles per
// cell less than a faces kernel followed by a cells kernel.
// Everything the thread reads is requested up front, before the first branch on a loaded value (cell count of the beam, end
// patches): one round trip to DRAM per thread instead of four or five dependent ones -- with one thread per cell there is no
// loop to hide them behind, only other warps.  Rows are clamped, not predicated: the arrays are allocated for the padded bundle.
#ifndef BEAM_UPDATE_MIN_BLOCKS
#define BEAM_UPDATE_MIN_BLOCKS 3
#endif
#ifndef BEAM_UPDATE_MIN_WARPS
#define BEAM_UPDATE_MIN_WARPS 12 // resident one-warp blocks per SM the register budget is cut for: 168 registers.  Measured at
                                 // 8192 beams (backward part, ms): 12 warps 0.301, 13 (152 registers) 0.315, 14 (144) 0.320, 16 (128) 0.320 -- spills
#endif
// A load the compiler must issue where it stands: left to itself it sinks every load below the early-exit branch and next to
// its first use (to save registers), which turns one round trip into a chain of them.
DEV double ldNow(const double* p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
DEV V3 ldNow3(const double* __restrict__ p, size_t i, size_t s) { return v3(ldNow(p + i), ldNow(p + i + s), ldNow(p + i + 2 * s)); }
// WARP_BLOCKS: one warp per block (a block = one cell row of one 32-beam tile).  The thousands of short-lived blocks then need
// no block barrier (the verdict of the previous iteration and the two partial sums stay inside the warp), a warp's registers
// are free again the moment IT is done rather than when the slowest of four is, and the look at the verdict comes after the
// loads have been requested instead of a round trip to L2 before them.
template <int SCHEME, bool WARP_BLOCKS>
__global__ void __launch_bounds__(WARP_BLOCKS ? 32 : 128, WARP_BLOCKS ? BEAM_UPDATE_MIN_WARPS : BEAM_UPDATE_MIN_BLOCKS)
beam_update(const __grid_constant__ BeamParams p, const double* __restrict__ X)
{
    if (!WARP_BLOCKS && convLeave(p.cv, true)) return;
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(t & 31), Rc = p.nmax, Rf = p.nmax + 1;
    const size_t row = t >> 5, T = row / (size_t)Rc, sb = p.Bpad;
    const int i = (int)(row - T * (size_t)Rc), b = (int)(T * 32) + lane;
    double v[2] = {0.0, 0.0};
    if (T * 32 < (size_t)p.Bpad) {
        const int iN = min(i + 1, Rc - 1);
        const size_t f3 = tileRow(T, Rf, i + 1, 3, lane), f4 = tileRow(T, Rf, i + 1, 4, lane);
        const size_t x0 = tileRow(T, Rc, i, 6, lane), xN = tileRow(T, Rc, iN, 6, lane);
        // from DRAM first ...
        const Q4 qfOld = q4(ldNow(p.Lf + f4), ldNow(p.Lf + f4 + TS), ldNow(p.Lf + f4 + 2 * TS), ldNow(p.Lf + f4 + 3 * TS));
        const V3 Kold = ldNow3(p.K, f3, TS), Qold = ldNow3(p.Q, f3, TS), Mold = ldNow3(p.M, f3, TS);
        const V3 WoldN = ldNow3(p.wSave, tileRow(T, Rc, iN, 3, lane), TS);
        CellOld old; // the cell's own rows (its W is still the old one: only this thread writes it)
        {
            const size_t c3 = tileRow(T, Rc, i, 3, lane), c4 = tileRow(T, Rc, i, 4, lane);
            old.W = ldNow3(p.W, c3, TS);
            old.q = q4(ldNow(p.Lam + c4), ldNow(p.Lam + c4 + TS), ldNow(p.Lam + c4 + 2 * TS), ldNow(p.Lam + c4 + 3 * TS));
            old.Th = ldNow3(p.Th, c3, TS);
            old.Ac = old.dOm = v3(0, 0, 0);
            if (SCHEME == BFK_NEWMARK) { old.Ac = ldNow3(p.Ac, c3, TS); old.dOm = ldNow3(p.dOm, c3, TS); }
        }
        const V3 DW = ldNow3(X, x0, TS), DT = ldNow3(X, x0 + 3 * TS, TS), xWn = ldNow3(X, xN, TS), xTn = ldNow3(X, xN + 3 * TS, TS);
        // ... then the per-beam constants (cache hits)
        const int n = p.nSeg[b];
        const double dZ = ldNow(p.dZ + b);
        const V3 CQ = ldNow3(p.CQ, b, sb), CM = ldNow3(p.CM, b, sb);
        M3 refL;
#pragma unroll
        for (int k = 0; k < 9; k++) refL.a[k] = ldNow(p.refL + b + k * sb);
        if (WARP_BLOCKS && p.cv.xbuf && p.cv.iter > 0 && convDecide(p.cv, p.cv.iter - 1, true) == 1) return; // the loop ended earlier
        if (b < p.B && i < n) {
            // ---- the cell
            const V3 Wold = old.W;
            v[0] = DW.x * DW.x + DW.y * DW.y + DW.z * DW.z + DT.x * DT.x + DT.y * DT.y + DT.z * DT.z;
            updateCellFrom<SCHEME>(p, T, lane, i, DW, DT, old, v[1]);
            // ---- the faces it owns
            const double dcI = 1.0 / dZ, dcB = 2.0 / dZ, delta = 1.0 / dcB;
            if (i < n - 1) { // internal face i+1 (Evolve.C:804-836, 881-913), owner = this cell, neighbour = cell i+1
                const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]), e0 = mulT(refL, dR0);
                const M3 LmOld = quatToMat(qfOld);
                const V3 tv = dR0 + dcI * (WoldN - Wold);
                const FaceCoef f = faceCoefficients(LmOld, CQ, CM, gammaOf(LmOld, e0, tv), Kold, Qold, Mold, tv, dcI, false);
                const FaceState o = updateFace(f, qfOld, LmOld, Kold, DW, xWn, DT, xTn, dcI, false);
                stq(p.Lf, f4, TS, o.Lf);
                st3(p.K, f3, TS, o.K); st3(p.Q, f3, TS, o.Q); st3(p.M, f3, TS, o.M);
            } else
                updatePatchFace(p, b, 1, n, CQ, CM, dcB, delta, Wold, DW, DT);
            if (i == 0) updatePatchFace(p, b, 0, n, CQ, CM, dcB, delta, Wold, DW, DT);
        }
    }
    if (WARP_BLOCKS) {
#pragma unroll
        for (int k = 0; k < 2; k++) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
            if (lane == 0) p.partial[(size_t)(1 + k) * p.nTiles + blockIdx.x] = v[k];
        }
    } else
        blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.x);
}
