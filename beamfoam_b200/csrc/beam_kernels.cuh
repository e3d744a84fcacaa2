// beam_kernels.cuh — the Newton-iteration kernels (FP64, sm_100a).
//
// One Newton iteration of coupledTotalLagNewtonRaphsonBeam::evolve() is two sweeps over every beam, launched as two
// kernels (beam_forward, beam_backward; beam_newton runs both in one launch and measures slower):
//
//   forwardSweep  : per beam, left -> right.  For every cell it recomputes the face
//                   coefficients (updateEqnCoefficients, Evolve.C:673-753), assembles the
//                   6x6 blocks d, l, u and the source (Matrix.C:55-475, BC.C:64-1190, loads and
//                   inertia Evolve.C:192-536) IN REGISTERS and immediately performs the
//                   block-Thomas forward elimination.  d/l/u/source never exist in HBM;
//                   only uPrime (36) and bPrime (6) per cell are written to scratch.
//   backwardSweep : per beam, right -> left.  Back-substitution, BC evaluate()
//                   (PF/*: momentBeamRotationNR, forceBeamDisplacementNR, followerForce,
//                   linearSpring), updateSolutionVariables (Evolve.C:757-923) and the partial
//                   sums of the three convergence norms (Evolve.C:588,618-626).
//
// Mapping: ONE THREAD PER BEAM, 32 beams per warp.  Cell, face and scratch arrays are
// "warp-tiled structure of arrays": the 32 beams of warp tile T = b/32 own one contiguous
// slab per array, element (cell i, component c, lane l = b%32) at
// (((T*R + i)*nc + c)*32 + l), R = rows per beam (nmax cells / nmax+1 faces).  The 32 lanes
// of a warp read 32 consecutive doubles (256 B, fully coalesced) for every scalar, a warp
// walks each array as ONE contiguous stream (a handful of 2 MB pages live per warp instead
// of one page per component row, which thrashed the TLB), and the serial block recurrence
// of each beam runs entirely in one thread's registers with no cross-lane traffic.
// Per-beam constants and end-patch arrays stay [c][Bpad].
#pragma once
#include "beam_math.cuh"
#include <stdint.h>

#define BFK_W_FIXED 0
#define BFK_W_FORCE 1
#define BFK_W_FOLLOWER 2
#define BFK_W_SPRING 3
#define BFK_TH_FIXED 0
#define BFK_TH_MOMENT 1
// d2dt2Schemes (the kernels are instantiated per scheme)
#define BFK_STEADY 0
#define BFK_NEWMARK 1
#define BFK_EULER 2

// The convergence test of the Newton loop (checkConvergence, Evolve.C:926-1026) evaluated ON THE DEVICE, so that
// bf_evolve can enqueue iterations ahead of the test: the sweeps of iteration k+1 start with a look at the global norms of
// iteration k and leave at once when the loop ended there.  Every rank owns an exchange buffer
// [2 banks][CONV_MAXK iterations][nRanks][4 doubles: sum source^2, sum dx^2, sum x^2, tag]; the 1-block reduction
// kernel of iteration k writes this rank's record into the buffer of EVERY rank (peer memory over NVLink, CUDA IPC) --
// a one-sided 32-byte store per peer instead of a collective -- and whoever needs the global sums adds the nRanks
// records in rank order (bit-identical on every rank).  tag = 2 * epoch of the evolve() call (+1: this rank had seen the
// loop end before iteration k: the record carries no sums).
#define CONV_MAXK 4096
#define CONV_MAX_RANKS 16
struct ConvCtl {
    const double* xbuf; // this rank's bank of the exchange buffer; null: the host decides (bf_newton_iteration, reducer callback)
    double* vflag;      // [CONV_MAXK] verdict cache of this bank: tag + 0.5 = the loop ended with (or before) iteration k, tag + 0.25 = it goes on
    int nRanks, iter;   // iter: the Newton iteration the launch belongs to
    double tag;
    double absTol, resTol, solTol, divTol;
    int nCorr;
    unsigned long long waitNs; // how long a kernel waits for a peer's record before it gives the peer up (BF_ERR_COMM)
};

struct BeamParams {
    ConvCtl cv;
    int B;        // beams in this context
    int Bpad;     // padded beam count (leading dimension)
    int nmax;     // max cells per beam
    int hL;       // split sweeps: the left part of a beam is cells 0 .. hL-1, the mirrored part cells hL .. nmax-1 (1 <= hL < nmax)
    int newmark;  // d2dt2Scheme: BFK_STEADY / BFK_NEWMARK / BFK_EULER (the kernels are instantiated per scheme)
    int first;    // 1 on the first Newton iteration of a time step (Newmark predictor)
    int storeInc; // keep DW/DTheta cell fields
    double dt, beta, gamma;
    // Newmark scalars, divided once on the host: 1/(beta dt), 0.5/beta - 1, gamma/(beta dt), 1/(beta dt^2), (1/dt)(gamma/beta)
    double nk1, nk2, nVel, nAcc, nU;
    // The arrays U and Om hold the Newmark step constants  U* = U - gamma dt Accl,  Omega* = Omega - gamma dt dotOmega
    // (= U_old + dt (1 - gamma) Accl_old, the explicit part of the Newmark velocity): every correction adds
    // gamma/(beta dt) DW to U and 1/(beta dt^2) DW to Accl (Evolve.C:782-789, 849-863), so U - gamma dt Accl does not
    // change inside a time step and neither U nor Omega has to be read or written by the corrections.
    // gdt = gamma * dt of this step; gdtPrev = gamma * dt of the step the stored constants belong to.
    double gdt, gdtPrev;
    // per beam [Bpad]
    const int* nSeg;
    const double* dZ;
    const double* CQ;     // [3][Bpad]
    const double* CM;     // [3][Bpad]
    const double* ARho;   // [Bpad]
    const double* CI;     // [3][Bpad]
    const double* weight; // [3][Bpad]
    const double* refL;   // [9][Bpad]
    const int* wType;     // [2][Bpad]
    const int* thType;    // [2][Bpad]
    // cells [tile][nmax][nc][32]; Lam = unit quaternion (4) of the cell rotation Lambda; U, Om = U*, Omega* (above)
    double *W, *Th, *Lam, *U, *Ac, *Om, *dOm, *DW, *DTh;
    const double *q, *m, *Lc; // optional (may be null)
    // optional point forces (Evolve.C:211-271), 9 per cell: sum F, sum zeta F over zeta > 0, sum zeta F over zeta < 0
    const double* pf;
    double* pfSrc; // [cells][6] source terms formed before each sweep by beam_point_forces and beam_momentum_contributions
                   // (beam_contrib.cuh), subtracted from the source; null when the case has neither
    // Euler d2dt2Scheme only: old-time level of W, U, Omega (3 each) and Lambda (quaternion), written by the first
    // iteration of a time step.  With Euler, U / Om hold the plain fields (no step constants) and dOm is not used.
    double *Wold, *Uold, *Omold, *Lamold;
    // faces [tile][nmax+1][nc][32]: face j of beam b; j=0 left patch, j=nSeg[b] right patch
    // Lf = unit quaternion (4) of the product Lambdaf & refLambdaf (the only form the interior of a beam
    // needs, Evolve.C:678,825-830,881-887).  Gamma is not stored: it is a pure function of the stored
    // W and Lambdaf (Evolve.C:881-887) and is recomputed where explicitQ needs it.
    double *Lf, *K, *Q, *M;
    // ends [2][nc][Bpad]
    double *Wb, *Thb, *Lamb, *DWb, *DThb, *force;
    double* ifW; // [3][Bpad] split sweeps: pre-update W of the cell left of the interface
    double* wSave; // cells [tile][nmax][3][32]: the pre-update W, saved by the forward part for beam_update (one thread per cell: a
                   // cell's thread reads its neighbour's OLD W while the neighbour's thread writes the new one); null otherwise
    const double *wPresc, *thPresc, *follower, *offset, *moment, *springK, *springDir;
    // block-Thomas scratch
    double* uP; // [tile][nmax][36][32]
    double* bP; // [tile][nmax][6][32]
    // per-block partial sums: partial[slot * nTiles + block], nTiles = the stride of the array (>= blocks of any launch)
    double* partial;
    int nTiles;
};

// Everything the sweeps write is next read a whole sweep later (3.7 GB of state + 4.4 GB of scratch against 126 MB of
// L2), so the stores are marked streaming (evict-first): they must not push the rows that the L2 prefetch / the TMA
// copies have staged for the next cells out of L2.
#ifdef BEAM_NO_STREAM_HINTS
DEV void stS(double* p, double v) { *p = v; }
#else
DEV void stS(double* p, double v) { __stcs(p, v); }
#endif
// scratch written by the forward launch, read by the backward launch: through L2 (BEAM_SCRATCH_L1: through L1 as well)
#ifdef BEAM_SCRATCH_L1
DEV double ldScratch(const double* p) { return *p; }
#else
DEV double ldScratch(const double* p) { return __ldcg(p); }
#endif
DEV V3 ld3(const double* __restrict__ p, size_t i, size_t s) { return v3(p[i], p[i + s], p[i + 2 * s]); }
DEV void st3(double* __restrict__ p, size_t i, size_t s, V3 v) { stS(p + i, v.x); stS(p + i + s, v.y); stS(p + i + 2 * s, v.z); }
DEV M3 ld9(const double* __restrict__ p, size_t i, size_t s)
{
    M3 r;
#pragma unroll
    for (int k = 0; k < 9; k++) r.a[k] = p[i + k * s];
    return r;
}
DEV void st9(double* __restrict__ p, size_t i, size_t s, const M3& A)
{
#pragma unroll
    for (int k = 0; k < 9; k++) p[i + k * s] = A.a[k];
}
DEV Q4 ldq(const double* __restrict__ p, size_t i, size_t s) { return q4(p[i], p[i + s], p[i + 2 * s], p[i + 3 * s]); }
DEV void stq(double* __restrict__ p, size_t i, size_t s, Q4 q) { stS(p + i, q.w); stS(p + i + s, q.x); stS(p + i + 2 * s, q.y); stS(p + i + 3 * s, q.z); }
// Gamma = refLambdaf.T() & ((Lambdaf.T() & dRdS) - dR0Ds)   Evolve.C:881-887, from the STORED W and
// Lm = Lambdaf & refLambdaf:  Gamma = Lm.T() & dRdS - e0,  e0 = refLambdaf.T() & dR0Ds
DEV V3 gammaOf(const M3& Lm, V3 e0, V3 t) { return mulT(Lm, t) - e0; }

// Warp-tiled addressing: TS = doubles between two components of one cell (the tile width);
// tileRow(T, R, r, nc, lane) = first component of row r (cell or face) of the lane's beam.
#define TS ((size_t)32)
DEV size_t tileRow(size_t T, int R, int r, int nc, int lane) { return ((T * (size_t)R + (size_t)r) * nc) * TS + lane; }

// L2 prefetch of the rows the NEXT loop iteration will read.  The serial recurrence of a beam
// leaves only ~2 resident warps per scheduler (register-bound), so the DRAM latency of the loads
// at the top of an iteration is exposed; pulling the lines into L2 one cell ahead turns it into an
// L2 hit.  Every lane prefetches its own element: the LSU coalesces a warp into two 128 B lines.
#if defined(BEAM_PREFETCH_L1)
DEV void pfL2(const double* __restrict__ p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#elif !defined(BEAM_NO_PREFETCH)
DEV void pfL2(const double* __restrict__ p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
DEV void pfL2(const double* __restrict__) {}
#endif
#ifdef BEAM_NEWMARK_PF_L1
DEV void pfNm(const double* __restrict__ p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#else
DEV void pfNm(const double* __restrict__ p) { pfL2(p); }
#endif
template <int NR>
DEV void pfRows(const double* __restrict__ p, size_t i0, size_t s)
{
#pragma unroll
    for (int k = 0; k < NR; k++) pfL2(p + i0 + (size_t)k * s);
}

// Coefficients of one face (updateEqnCoefficients, Evolve.C:673-753), from the PRE-solve state.
struct FaceCoef {
    M3 CQW, CQTheta, CMTheta, CMQW, CMQTheta; // CQDTheta == 0 (CTL.C:586-603), CMTheta2 = -spin(M)
    V3 eQ, eM, eMQ, Mold, t;
};

// dc = deltaCoeffs of the face; boundary faces get the x2 of Evolve.C:744-752.
// Lm = Lambdaf & refLambdaf (:678): the device stores this product (as a quaternion), not Lambdaf.
DEV FaceCoef faceCoefficients(const M3& Lm, V3 CQ, V3 CM, V3 Gam, V3 K, V3 Q, V3 M, V3 t,
                              double dc, bool boundary)
{
    FaceCoef f;
    f.CQW = conjDiag(Lm, CQ);                  // :680
    f.eQ = mul(Lm, cmul(CQ, Gam));             // :682
    f.eM = mul(Lm, cmul(CM, K));               // :684
    const M3 SQ = spin(Q);
    f.CQTheta = mulSpinR(f.CQW, t) - SQ;       // :686-693
    f.CMTheta = conjDiag(Lm, CM);              // :699
    f.Mold = M;                                // CMTheta2 = -spin(M) :703
    const M3 SC = mulSpinL(t, f.CQW);
    const double h = boundary ? 1.0 / dc : 0.5 / dc; // 0.5*(..)/deltaCoeffs, x2 on patches
    f.CMQW = h * (SC - SQ);                    // :706-715
    f.CMQTheta = h * (mulSpinR(SC, t) - mulSpinL(t, SQ)); // :718-735
    f.eMQ = h * cross(t, f.eQ);                // :737-741
    f.t = t;
    return f;
}

DEV void addBlk(double (&D)[6][6], int r0, int c0, double s, const M3& T)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) D[r0 + r][c0 + c] += s * T.a[3 * r + c];
}
DEV void addBlkS(double (&D)[6][6], int r0, int c0, double s, const Sym3& T) { addBlk(D, r0, c0, s, full(T)); }
// D[r0.., c0..] += s * T.T()
DEV void addBlkT(double (&D)[6][6], int r0, int c0, double s, const M3& T)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) D[r0 + r][c0 + c] += s * T.a[3 * c + r];
}
// D[r0.., c0..] = carry block from the stash + s * T  (rows k0.. of the warp slab hold the carry, 6x6 row-major)
DEV void carryBlkFrom(double (&D)[6][6], const double* sm, int k0, int r0, int c0, double s, const M3& T)
{
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
            D[r0 + r][c0 + c] = ((const volatile double*)sm)[(k0 + 6 * (r0 + r) + c0 + c) * 32] + s * T.a[3 * r + c];
}
DEV void subV(double (&b)[6], int r0, V3 v) { b[r0] -= v.x; b[r0 + 1] -= v.y; b[r0 + 2] -= v.z; }
DEV void addV(double (&b)[6], int r0, V3 v) { b[r0] += v.x; b[r0 + 1] += v.y; b[r0 + 2] += v.z; }

// End-patch values a thread needs for one end of its beam.
struct EndBC {
    int wType, thType;
    V3 Wb, Thb;        // stored patch values (== prevIter at iteration start)
    V3 wPresc, thPresc;
    V3 force;          // force() as seen by the assembly of this iteration
    V3 moment;
};

DEV EndBC loadEnd(const BeamParams& p, int end, int b, const M3& Lfb)
{
    const size_t s = p.Bpad;
    EndBC e;
    e.wType = p.wType[end * s + b];
    e.thType = p.thType[end * s + b];
    const size_t i3 = (size_t)end * 3 * s + b;
    e.Wb = ld3(p.Wb, i3, s);
    e.Thb = ld3(p.Thb, i3, s);
    e.wPresc = ld3(p.wPresc, i3, s);
    e.thPresc = ld3(p.thPresc, i3, s);
    e.moment = ld3(p.moment, i3, s);
    if (e.wType == BFK_W_FOLLOWER) // followerForce...C:183-265: force = Lambdaf_b . F_material, every iteration
        e.force = mul(Lfb, ld3(p.follower, i3, s));
    else
        e.force = ld3(p.force, i3, s);
    return e;
}

// assembleBoundaryConditions for ONE end patch (BC.C:64-1190) into the diagonal block D and
// the source b of its face cell.  pDelta = 1/deltaCoeffs_b.
DEV void bcClosure(double (&D)[6][6], double (&b)[6], const FaceCoef& f, const EndBC& e, const M3& Lfb,
                   double pDelta)
{
    const double ipd = 1.0 / pDelta;
    const bool isForce = e.wType != BFK_W_FIXED;
    const bool isFollower = e.wType == BFK_W_FOLLOWER;
    const bool isMoment = e.thType == BFK_TH_MOMENT;
    const M3 CMTheta2 = -1.0 * spin(f.Mold);
    // Lambdaf_b & (pTheta - pThetaPrev).  Only a fixed-Theta patch changes its value in
    // updateCoeffs(); for a moment patch pTheta == pThetaPrev (storePrevIter, Evolve.C:764-765).
    const V3 pThetaCorr = isMoment ? v3(0, 0, 0) : mul(Lfb, e.thPresc - e.Thb);
    const V3 pWCorr = e.wPresc - e.Wb;                 // pW - pWPrev
    M3 invCM, Bm;
    V3 A = v3(0, 0, 0);
    if (isMoment) {
        invCM = inv(ipd * f.CMTheta + CMTheta2);
        A = mul(invCM, e.moment - f.eM);
        Bm = mul(invCM, ipd * f.CMTheta);
    }
    // ---- W equation
    if (isForce) { // :102-155
        subV(b, 0, e.force);
        addV(b, 0, f.eQ);
        if (isFollower) addBlk(D, 0, 3, -1.0, spin(e.force)); // :128-154
    } else { // :339-521
        addBlk(D, 0, 0, -ipd, f.CQW);
        const V3 wc = v3(pWCorr.x + BF_SMALL, pWCorr.y + BF_SMALL, pWCorr.z + BF_SMALL); // :386-393
        subV(b, 0, ipd * mul(f.CQW, wc));
        if (isMoment) { // :400-458
            addBlk(D, 0, 3, 1.0, mul(f.CQTheta, Bm));
            subV(b, 0, mul(f.CQTheta, A));
        } else { // :459-520 (CQDTheta == 0)
            subV(b, 0, mul(f.CQTheta, pThetaCorr));
        }
    }
    subV(b, 0, f.eQ); // :524-529
    // ---- Theta equation
    if (isMoment) { // :534-583
        subV(b, 3, e.moment - f.eM);
    } else { // :584-677
        addBlk(D, 3, 3, -ipd, f.CMTheta);
        subV(b, 3, ipd * mul(f.CMTheta, pThetaCorr));
        subV(b, 3, mul(CMTheta2, pThetaCorr));
    }
    subV(b, 3, f.eM); // :680-685
    // ---- dr x Q term
    if (isForce) { // :694-878
        subV(b, 3, pDelta * cross(f.t, e.force - f.eQ)); // :714-730
        M3 iCC;
        if (isMoment || !isFollower) iCC = mul(inv(f.CQW), f.CQTheta);
        if (isMoment) { // :732-791
            subV(b, 3, mul(f.CMQTheta, A) - mul(f.CMQW, mul(iCC, A)));
            addBlk(D, 3, 3, 1.0, mul(f.CMQTheta, Bm) - mul(f.CMQW, mul(iCC, Bm)));
        }
        if (isFollower) { // :793-835
            addBlk(D, 3, 3, -pDelta, mulSpinL(f.t, spin(e.force)));
        } else { // :836-877 — runs for every Theta class (all derive from fixedValue)
            const V3 dwds1 = -mul(iCC, pThetaCorr);
            subV(b, 3, mul(f.CMQW, dwds1) + mul(f.CMQTheta, pThetaCorr));
        }
    } else { // :1034-1175
        addBlk(D, 3, 0, -ipd, f.CMQW);
        const V3 wc = v3(pWCorr.x + BF_SMALL, pWCorr.y + BF_SMALL, pWCorr.z + BF_SMALL);
        subV(b, 3, ipd * mul(f.CMQW, wc));
        if (isMoment) { // :1089-1138
            addBlk(D, 3, 3, 1.0, mul(f.CMQTheta, Bm));
            subV(b, 3, mul(f.CMQTheta, A));
        } else { // :1139-1174
            subV(b, 3, mul(f.CMQTheta, pThetaCorr));
        }
    }
    subV(b, 3, f.eMQ); // :1178-1187
}

DEV double ldAcquireSys(const double* p)
{
    double v;
    asm volatile("ld.acquire.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
DEV void stReleaseSys(double* p, double v) { asm volatile("st.release.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
// Global sums of iteration k, added in rank order, by ONE WARP: lane r reads the record of rank r (a system-scope load is a
// round trip of the order of a microsecond; one thread reading 4 nRanks of them in turn cost 45 us per block on 8 GPUs), the
// sums are then added in rank order through shuffles (bit-identical on every rank).  wait: spin until every rank's record is
// there.  s[3] = the sums of iteration 0 (the initial residual; its records arrived long ago).  Returns -1: a record is missing
// (only without wait), 1: a rank had seen the loop end before iteration k, 2: a peer's record did not arrive within
// ConvCtl::waitNs (the peer died or never launched: the kernels must not spin on the GPU for ever -- the loop is treated as ended
// and the host reports BF_ERR_COMM), 0: s holds the sums.  Call with all 32 lanes.
DEV unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
DEV int convSums(const ConvCtl& c, const int k, const bool wait, double (&s)[4])
{
    const int lane = threadIdx.x & 31;
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, i0 = 0.0;
    int state = 0;
    if (lane < c.nRanks) {
        const double* rec = c.xbuf + ((size_t)k * c.nRanks + lane) * 4;
        double tag;
        unsigned long long t0 = 0;
        for (unsigned spins = 0;; spins++) {
            tag = ldAcquireSys(rec + 3);
            if (tag == c.tag || tag == c.tag + 1.0) break;
            if (!wait) { state = -1; break; }
            __nanosleep(100);
            if ((spins & 1023u) == 1023u) { // look at the clock every ~0.1 ms of waiting
                const unsigned long long t = globalTimerNs();
                if (!t0) t0 = t;
                else if (t - t0 > c.waitNs) { state = 2; break; }
            }
        }
        if (state == 0) {
            if (tag != c.tag) state = 1;
            v0 = ldAcquireSys(rec); v1 = ldAcquireSys(rec + 1); v2 = ldAcquireSys(rec + 2);
            i0 = k > 0 ? ldAcquireSys(c.xbuf + (size_t)lane * 4) : v0;
        }
    }
    __syncwarp();
    const bool missing = __any_sync(0xffffffffu, state == -1), ended = __any_sync(0xffffffffu, state == 1);
    const bool lost = __any_sync(0xffffffffu, state == 2);
    s[0] = s[1] = s[2] = s[3] = 0.0;
    for (int r = 0; r < c.nRanks; r++) {
        s[0] += __shfl_sync(0xffffffffu, v0, r); s[1] += __shfl_sync(0xffffffffu, v1, r);
        s[2] += __shfl_sync(0xffffffffu, v2, r); s[3] += __shfl_sync(0xffffffffu, i0, r);
    }
    return lost ? 2 : missing ? -1 : ended ? 1 : 0;
}
// The five exit rules of checkConvergence (Evolve.C:926-1026) on the squared global sums: does the loop end with iteration k?
DEV bool convRules(const ConvCtl& c, const int k, const double (&s)[4])
{
    if (!(s[0] == s[0]) || !(s[1] == s[1]) || !(s[2] == s[2])) return true; // non-finite: a singular block
    const double cur = sqrt(s[0]), dx = sqrt(s[1]), xn = sqrt(s[2]);
    const double init = k == 0 ? cur : sqrt(s[3]);
    if (cur <= c.absTol) return true;
    if (cur <= c.resTol * init) return true;
    if (dx <= c.solTol * xn) return true;
    if (cur >= c.divTol * init && k > 1) return true;
    return k >= c.nCorr;
}
// 1: the loop ends with iteration k or had ended before, 0: it goes on, -1: not known yet (only without wait).  The verdict is
// the same for whoever computes it: the first warp that has all the records caches it (ConvCtl::vflag), and the thousands of
// short-lived blocks of the cell-parallel kernels then pay one L2 load.  Call with all 32 lanes of a warp.
DEV int convDecide(const ConvCtl& c, const int k, const bool wait)
{
    volatile double* vf = c.vflag + k;
    const double v = *vf;
    if (v == c.tag + 0.5) return 1;
    if (v == c.tag + 0.25) return 0;
    double s[4];
    const int m = convSums(c, k, wait, s);
    if (m < 0) return -1;
    const int d = m ? 1 : convRules(c, k, s) ? 1 : 0; // a lost peer (m == 2) ends the loop on this rank as well
    if ((threadIdx.x & 31) == 0) *vf = c.tag + (d == 1 ? 0.5 : 0.25);
    return d;
}
// Prologue of every sweep kernel: true = leave, the loop ended with an earlier iteration.  The forward sweeps do not wait
// (if the records of the previous iteration are not all there yet they run speculatively: they write scratch and partial
// sums only, and the time they take is time this GPU would have waited for the slowest one); the backward sweeps, which
// change the state, wait.
DEV bool convLeave(const ConvCtl& c, const bool wait)
{
    if (!c.xbuf || c.iter == 0) return false;
    __shared__ int sLeave;
    if (threadIdx.x < 32) {
        const int d = convDecide(c, c.iter - 1, wait);
        if (threadIdx.x == 0) sLeave = d == 1;
    }
    __syncthreads();
    return sLeave != 0;
}

// Block reduction of up to 3 values; thread 0 writes partial[(slot0 + k)*nTiles + tile].
template <int NV>
DEV void blockReduceStore(double (&v)[NV], double* __restrict__ partial, int slot0, int nTiles, int tile)
{
    __shared__ double sm[NV][32];
    __syncthreads(); // the staging words may still be read by the previous reduction of a persistent block
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) sm[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
        const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
        for (int k = 0; k < NV; k++) {
            double x = lane < nw ? sm[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) partial[(size_t)(slot0 + k) * nTiles + tile] = x;
        }
    }
}

// ---------------------------------------------------------------------------
// forward: coefficients + assembly + block-Thomas elimination
// ---------------------------------------------------------------------------
// Every warp owns a private slab of shared memory, [row][32 lanes] doubles (row r of lane l at
// sm[r*32 + l]: consecutive lanes hit consecutive banks, every access is conflict-free):
//   * per-beam constants (12 rows; CIRho and the self-weight are used once per cell and come from L1),
//   * the "stash": the face tensors of face i+1 that are needed again after the elimination of
//     cell i (u columns, l block, neighbour part of d_{i+1}); a register extension -- an FP64 warp
//     instruction occupies the pipe for 2 cycles, an LDS.64 costs the same,
//   * the prefetch buffer: the face rows the NEXT cell will read first, brought in by the TMA
//     engine (cp.async.bulk, one copy per array, completion on the warp's mbarrier) while the
//     current cell is being eliminated.  The warp-tiled layout makes every array's rows of one
//     cell one contiguous run (nc x 256 B).  The Newmark cell rows are consumed late in the
//     assembly and are loaded directly (L2-prefetched one cell ahead),
//   * the carry of the block-Thomas recurrence (36 doubles): written at the end of a cell, read
//     back where the next cell first needs each block, so that the 72 registers it would occupy
//     across the loop are free for the assembly.
#ifndef BEAM_THREADS
#define BEAM_THREADS 128
#endif
#ifndef BEAM_PF_DIST
#define BEAM_PF_DIST 1 // cells of look-ahead of the backward sweep's L2 prefetch
#endif
#ifndef BEAM_PFCELL_DIST
#define BEAM_PFCELL_DIST 1 // the L2 prefetch of the Newmark cell rows runs this many cells ahead of their cp.async
#endif
#ifndef BEAM_MIN_BLOCKS
#define BEAM_MIN_BLOCKS 2
#endif
#ifndef BEAM_BWD_MIN_BLOCKS
#define BEAM_BWD_MIN_BLOCKS BEAM_MIN_BLOCKS
#endif
enum {
    R_DR0 = 0,   // dR0Ds = refLambdaf & e_x (3)
    R_DZ = 3,    // dZ of the beam (1)
    R_ARHO = 4,  // rho A (1)
    R_N = 5,     // cells of the beam, as a double (1).  These three are read once per cell: kept here they cannot
                 // be spilled to local memory, whose reloads go to L2 (L1 is all but carved out for shared memory)
    R_CQ = 6,    // diag CQ (3)
    R_CM = 9,    // diag CM (3)
    R_LB = 12,   // -(l_{i-1} & bPrime_{i-1}) (6): the right-hand-side part of the carry
    R_P = 18,    // a*CQW, symmetric: xx xy xz yy yz zz (6)   face i+1
    R_QT = 24,   // 0.5*CQTheta (9);  a*CMQW = -(0.5*CQTheta).T()
    R_S1 = 33,   // a*CMTheta, symmetric (6)
    R_S3 = 39,   // 0.5*CMQTheta (9)
    R_MF = 48,   // 0.5 M of face i+1 (3): 0.5*CMTheta2 = -spin(0.5 M)
    R_SC = 51,   // neighbour part of the source of cell i+1 (6)
    R_BUF = 57,  // prefetch buffer
    B_W = 0,     //   W of cell i+1 (3)
    B_LF = 3,    //   Lambdaf & refLambdaf of face i+1, quaternion (4)
    B_K = 7, B_Q = 10, B_M = 13, //   K, Q, M of face i+1
    B_ROWS = 16,
    // second use of the same buffer inside a cell (Newmark): the rows of cell i that the inertia terms read, brought in
    // by the TMA engine while the face part runs (completion on the warp's second mbarrier)
    C_LAM = 0,   //   Lambda of cell i, quaternion (4)
    C_AC = 4, C_OM = 7, C_DOM = 10, //   Accl, Omega*, dotOmega of cell i
    C_U = 13,    //   U* of cell i (only read on the first iteration of a time step)
    R_DC = R_BUF + B_ROWS, // the carry: (nei part of d_i) - l_{i-1} & uPrime_{i-1}, row-major 6x6 (36)
    R_CI = R_DC + 36,      // diag CIRho (3): a global load per cell was an exposed L2 round trip (L1 is carved out for this slab)
    R_TOTAL = R_CI + 3     // 112 rows = 28 KB per warp: two 4-warp blocks per SM fill the 228 KB exactly
};
#define WARP_SMEM_DOUBLES (R_TOTAL * 32)
#define SMEM_HEADER_BYTES 128 // one mbarrier (8 B) per warp
#define SMEM_BYTES (SMEM_HEADER_BYTES + (BEAM_THREADS / 32) * WARP_SMEM_DOUBLES * sizeof(double))

#define carryBlk(D, sm, r0, c0, s, T) carryBlkFrom(D, sm, R_DC, r0, c0, s, T)
DEV double smLd(const double* sm, int k) { return ((const volatile double*)sm)[k * 32]; }
DEV void smSt(double* sm, int k, double v) { sm[k * 32] = v; }
DEV V3 smLd3(const double* sm, int k) { return v3(smLd(sm, k), smLd(sm, k + 1), smLd(sm, k + 2)); }
DEV void smSt3(double* sm, int k, V3 v) { smSt(sm, k, v.x); smSt(sm, k + 1, v.y); smSt(sm, k + 2, v.z); }
DEV Q4 smLdq(const double* sm, int k) { return q4(smLd(sm, k), smLd(sm, k + 1), smLd(sm, k + 2), smLd(sm, k + 3)); }
DEV M3 smLd9(const double* sm, int k)
{
    M3 r;
#pragma unroll
    for (int j = 0; j < 9; j++) r.a[j] = smLd(sm, k + j);
    return r;
}
DEV void smSt9(double* sm, int k, const M3& A)
{
#pragma unroll
    for (int j = 0; j < 9; j++) smSt(sm, k + j, A.a[j]);
}
DEV Sym3 smLdS(const double* sm, int k)
{
    Sym3 r;
    r.xx = smLd(sm, k); r.xy = smLd(sm, k + 1); r.xz = smLd(sm, k + 2);
    r.yy = smLd(sm, k + 3); r.yz = smLd(sm, k + 4); r.zz = smLd(sm, k + 5);
    return r;
}
DEV void smStS(double* sm, int k, const Sym3& A)
{
    smSt(sm, k, A.xx); smSt(sm, k + 1, A.xy); smSt(sm, k + 2, A.xz);
    smSt(sm, k + 3, A.yy); smSt(sm, k + 4, A.yz); smSt(sm, k + 5, A.zz);
}

// ---- mbarrier + bulk async copy (TMA engine, linear mode: no tensor map needed) ----
DEV uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbarInit(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
DEV void mbarWait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "BEAM_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra BEAM_MBAR_DONE;\n"
        "bra BEAM_MBAR_WAIT;\n"
        "BEAM_MBAR_DONE:\n"
        "}" ::"r"(smemAddr(bar)), "r"(parity) : "memory");
}
DEV void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// Per-lane asynchronous copies global -> shared (LDGSTS): no register holds the value in flight, and a lane only ever
// touches its own column of the slab, so no warp-level synchronisation is needed around them.
DEV void cpAsync8(double* dstSmem, const double* srcGlobal)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smemAddr(dstSmem)), "l"(srcGlobal) : "memory");
}
DEV void cpAsyncWaitAll() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Orders this thread's generic-proxy accesses to shared memory (the LDS reads of the buffer, the cp.async writes into it)
// before later async-proxy writes to the same rows (the next bulk copy).  Every lane executes it, then the warp
// synchronises, then the elected lane issues the copy.
#ifdef BEAM_NO_PROXY_FENCE
DEV void fenceProxyAsync() {}
#else
DEV void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// The warp's 16-row buffer is filled twice per cell:
//   fwdPrefetchFace(i): face i+1 and W of cell i+1, by the TMA engine (one elected lane, one bulk copy per array,
//                       completion on the warp's mbarrier) -- issued in the middle of cell i-1, waited for at the top
//                       of cell i: from DRAM, with two thirds of a cell to cover it;
//   fwdPrefetchCell(i): the Newmark rows of cell i, by per-lane cp.async from L2 (they are L2-prefetched half a cell
//                       earlier) -- issued at the top of cell i as soon as the face rows are in registers, waited for
//                       after the face part.  A bulk copy is too slow for this window (its completion was 0.9 us late
//                       per cell), and plain loads "issued early" do not work either: at 255 registers ptxas sinks them
//                       to their first use and the warp sits out an L2 round trip per cell (17 % of the sweep).
DEV int laneId()
{
    int l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l)); // re-read where needed: a value kept across the cell gets spilled
    return l;
}
// cellRow: the cell whose Newmark / Euler rows are pulled into L2 (its cp.async follows a phase later); wRow, faceRow: the
// rows of W and of the face arrays that the bulk copies fetch (the neighbour cell and the face towards it).  A
// left-to-right sweep passes (i, i + 1, i + 1); the mirrored half of a split beam walks the rows downwards.
template <int SCHEME>
DEV void fwdPrefetchFaceRows(const BeamParams& p, size_t T, double* warpSm, uint64_t* bar, int cellRow, int wRow, int faceRow)
{
    const int Rc = p.nmax, Rf = p.nmax + 1;
    const int lane = laneId();
    if (SCHEME == BFK_EULER) { // the cell rows the Euler inertia terms read (fwdPrefetchCell)
        const int ic = max(min(cellRow, Rc - 1), 0);
        const size_t n3 = tileRow(T, Rc, ic, 3, 0) + 3 * lane;
        pfNm(p.Lam + tileRow(T, Rc, ic, 4, 0) + 4 * lane);
        pfNm(p.U + n3);
        pfNm(p.Om + n3);
        if (!p.first) { pfNm(p.Uold + n3); pfNm(p.Omold + n3); }
    }
    if (SCHEME == BFK_NEWMARK) { // the cell rows of the same cell: pull them into L2 now, the cp.async follows a phase later
        // one address per lane and array: nc rows x 32 doubles are contiguous, lane l touches double nc*l, so every
        // 32-byte sector of the nc x 256 B run is requested once
        const int ic = max(min(cellRow, Rc - 1), 0);
        const size_t n3 = tileRow(T, Rc, ic, 3, 0) + 3 * lane;
        pfNm(p.Lam + tileRow(T, Rc, ic, 4, 0) + 4 * lane);
        pfNm(p.Ac + n3);
        pfNm(p.Om + n3);
        pfNm(p.dOm + n3);
        if (p.first) pfNm(p.U + n3);
    }
    if (lane != 0) return;
    const bool hasW = wRow >= 0 && wRow < p.nmax; // no neighbour cell in the slab past either end
    mbarExpectTx(bar, ((hasW ? 3 : 0) + 4 + 9) * 256u);
    double* buf = warpSm + (size_t)R_BUF * 32;
    if (hasW) bulkLoad(buf + B_W * 32, p.W + tileRow(T, Rc, wRow, 3, 0), 3 * 256, bar);
    bulkLoad(buf + B_LF * 32, p.Lf + tileRow(T, Rf, faceRow, 4, 0), 4 * 256, bar);
    bulkLoad(buf + B_K * 32, p.K + tileRow(T, Rf, faceRow, 3, 0), 3 * 256, bar);
    bulkLoad(buf + B_Q * 32, p.Q + tileRow(T, Rf, faceRow, 3, 0), 3 * 256, bar);
    bulkLoad(buf + B_M * 32, p.M + tileRow(T, Rf, faceRow, 3, 0), 3 * 256, bar);
}
template <int SCHEME>
DEV void fwdPrefetchFace(const BeamParams& p, size_t T, double* warpSm, uint64_t* bar, int i)
{
    fwdPrefetchFaceRows<SCHEME>(p, T, warpSm, bar, i + BEAM_PFCELL_DIST - 1, i + 1, i + 1);
}
template <int SCHEME>
DEV void fwdPrefetchCell(const BeamParams& p, size_t T, double* sm, int i)
{
    const int Rc = p.nmax, lane = laneId();
    const size_t c3 = tileRow(T, Rc, i, 3, lane), c4 = tileRow(T, Rc, i, 4, lane);
    double* buf = sm + (size_t)R_BUF * 32;
#pragma unroll
    for (int k = 0; k < 4; k++) cpAsync8(buf + (C_LAM + k) * 32, p.Lam + c4 + k * TS);
    if (SCHEME == BFK_NEWMARK) {
#pragma unroll
        for (int k = 0; k < 3; k++) {
            cpAsync8(buf + (C_AC + k) * 32, p.Ac + c3 + k * TS);
            cpAsync8(buf + (C_OM + k) * 32, p.Om + c3 + k * TS);
            cpAsync8(buf + (C_DOM + k) * 32, p.dOm + c3 + k * TS);
        }
        if (p.first) {
#pragma unroll
            for (int k = 0; k < 3; k++) cpAsync8(buf + (C_U + k) * 32, p.U + c3 + k * TS);
        }
    } else { // Euler: U, Omega and their old-time level (in the rows Newmark uses for Accl and dotOmega)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            cpAsync8(buf + (C_U + k) * 32, p.U + c3 + k * TS);
            cpAsync8(buf + (C_OM + k) * 32, p.Om + c3 + k * TS);
        }
        if (!p.first) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
                cpAsync8(buf + (C_AC + k) * 32, p.Uold + c3 + k * TS);
                cpAsync8(buf + (C_DOM + k) * 32, p.Omold + c3 + k * TS);
            }
        }
    }
}

// Boundary closure is executed for 2 of the n cells of a beam; keeping it out of line keeps its
// temporaries out of the register allocation of the steady-state loop.
__device__ __noinline__ void patchClosure(const BeamParams& p, int end, int b, V3 Wc, double* D36, double* src6)
{
    const size_t s = p.Bpad;
    const double dZ = p.dZ[b];
    const double dcB = 2.0 / dZ, pDelta = 1.0 / dcB;
    const M3 refL = ld9(p.refL, b, s);
    const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
    const int j = end ? p.nSeg[b] : 0;
    const size_t f3 = tileRow(b >> 5, p.nmax + 1, j, 3, b & 31), f4 = tileRow(b >> 5, p.nmax + 1, j, 4, b & 31);
    const M3 Lmb = quatToMat(ldq(p.Lf, f4, TS));  // Lambdaf_b & refLambdaf
    const M3 Lfb = mulBT(Lmb, refL);              // Lambdaf_b
    const EndBC e = loadEnd(p, end, b, Lfb);
    const V3 Wb = e.wType == BFK_W_FIXED ? e.wPresc : e.Wb; // after updateCoeffs()
    // left patch: outward normal -x => dR0Ds = -dR0 (CTL.C:987-994)
    const V3 dR0s = end ? dR0 : -dR0;
    const V3 t = dR0s + dcB * (Wb - Wc);
    const V3 Gam = gammaOf(Lmb, mulT(refL, dR0s), dR0s + dcB * (e.Wb - Wc)); // Gamma of the previous update: stored patch W
    const FaceCoef f = faceCoefficients(Lmb, ld3(p.CQ, b, s), ld3(p.CM, b, s), Gam,
                                        ld3(p.K, f3, TS), ld3(p.Q, f3, TS), ld3(p.M, f3, TS), t, dcB, true);
    double D[6][6], src[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        src[r] = src6[r];
#pragma unroll
        for (int c = 0; c < 6; c++) D[r][c] = D36[6 * r + c];
    }
    bcClosure(D, src, f, e, Lfb, pDelta);
#pragma unroll
    for (int r = 0; r < 6; r++) {
        src6[r] = src[r];
#pragma unroll
        for (int c = 0; c < 6; c++) D36[6 * r + c] = D[r][c];
    }
}

// What a cell reads from the prefetch buffer: copied to registers first, so that the buffer can be
// refilled for the next cell before any of the heavy work starts.
struct FaceIn {
    V3 Wn;     // W of cell i+1
    Q4 qf;     // Lambdaf & refLambdaf of face i+1
    V3 K, Q, M;
};
DEV FaceIn loadFaceIn(const double* sm)
{
    FaceIn f;
    f.Wn = smLd3(sm, R_BUF + B_W);
    f.qf = smLdq(sm, R_BUF + B_LF);
    f.K = smLd3(sm, R_BUF + B_K);
    f.Q = smLd3(sm, R_BUF + B_Q);
    f.M = smLd3(sm, R_BUF + B_M);
    return f;
}

// Point forces on cell i (Evolve.C:211-271): (sum F0 ; sum S(DR) F0) with DR = zeta * (dRdS of the face on that side
// of the cell) * (0.5/deltaCoeffs on an internal face, 1/deltaCoeffs_b on an end patch); dRdS from the pre-solve W,
// patch W after updateCoeffs().  Runs in its own small kernel (beam_point_forces) ahead of the sweeps, so that a case
// without constant/pointForces carries none of this in the hot loop.
DEV void pointForceSource(const BeamParams& p, const int b, const int i, double* out6)
{
    const size_t T = (size_t)(b >> 5), sb = p.Bpad;
    const int lane = b & 31, n = p.nSeg[b];
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ;
    const V3 dR0 = v3(p.refL[b], p.refL[3 * sb + b], p.refL[6 * sb + b]);
    const size_t f9 = tileRow(T, p.nmax, i, 9, lane);
    const V3 F0 = ld3(p.pf, f9, TS), Zp = ld3(p.pf, f9 + 3 * TS, TS), Zn = ld3(p.pf, f9 + 6 * TS, TS);
    const V3 Wc = ld3(p.W, tileRow(T, p.nmax, i, 3, lane), TS);
    auto patchW = [&](int end) {
        const size_t i3 = (size_t)end * 3 * sb + b;
        return p.wType[end * sb + b] == BFK_W_FIXED ? ld3(p.wPresc, i3, sb) : ld3(p.Wb, i3, sb);
    };
    V3 uP, uN;
    if (i < n - 1) uP = (0.5 * dZ) * (dR0 + dcI * (ld3(p.W, tileRow(T, p.nmax, i + 1, 3, lane), TS) - Wc));
    else uP = (1.0 / dcB) * (dR0 + dcB * (patchW(1) - Wc));
    if (i > 0) uN = (0.5 * dZ) * (dR0 + dcI * (Wc - ld3(p.W, tileRow(T, p.nmax, i - 1, 3, lane), TS)));
    else uN = (1.0 / dcB) * (dcB * (patchW(0) - Wc) - dR0);
    const V3 M0 = cross(uP, Zp) + cross(uN, Zn);
    out6[0] = F0.x; out6[1] = F0.y; out6[2] = F0.z; out6[3] = M0.x; out6[4] = M0.y; out6[5] = M0.z;
}

// Loads (Evolve.C:193-279) and inertia (Evolve.C:320-536, with the Newmark predictor :167-186 on the first iteration
// of a time step) of cell i, added to its diagonal block and source.
// The rows of a cell that the inertia terms read: the forward sweep brings them in by cp.async (fwdPrefetchCell /
// loadCellInSm); the cell-parallel path loads them directly (loadCellIn).
struct CellIn {
    Q4 ql;             // Lambda of the cell
    V3 Ac, Om, dOm, U; // Newmark: Accl, Omega*, dotOmega, U* (U* on the first iteration only)
                       // Euler:   U.oldTime, Omega, Omega.oldTime, U
};
template <int SCHEME>
DEV CellIn loadCellInSm(const double* sm, bool first)
{
    CellIn c;
    c.ql = smLdq(sm, R_BUF + C_LAM);
    c.Om = smLd3(sm, R_BUF + C_OM);
    if (SCHEME == BFK_NEWMARK) {
        c.Ac = smLd3(sm, R_BUF + C_AC);
        c.dOm = smLd3(sm, R_BUF + C_DOM);
        c.U = first ? smLd3(sm, R_BUF + C_U) : v3(0, 0, 0);
    } else { // on the first iteration of a step the old-time level IS the current one (the backward sweep stores it)
        c.U = smLd3(sm, R_BUF + C_U);
        c.Ac = first ? c.U : smLd3(sm, R_BUF + C_AC);
        c.dOm = first ? c.Om : smLd3(sm, R_BUF + C_DOM);
    }
    return c;
}
template <int SCHEME>
DEV CellIn loadCellIn(const BeamParams& p, const int b, const int i)
{
    CellIn c;
    c.ql = q4(1, 0, 0, 0);
    c.Ac = c.Om = c.dOm = c.U = v3(0, 0, 0);
    const size_t c3 = tileRow((size_t)(b >> 5), p.nmax, i, 3, b & 31), c4 = tileRow((size_t)(b >> 5), p.nmax, i, 4, b & 31);
    if (SCHEME == BFK_NEWMARK) {
        c.ql = ldq(p.Lam, c4, TS); c.Ac = ld3(p.Ac, c3, TS); c.Om = ld3(p.Om, c3, TS); c.dOm = ld3(p.dOm, c3, TS);
        if (p.first) c.U = ld3(p.U, c3, TS);
    }
    if (SCHEME == BFK_EULER) {
        c.ql = ldq(p.Lam, c4, TS); c.U = ld3(p.U, c3, TS); c.Om = ld3(p.Om, c3, TS);
        c.Ac = p.first ? c.U : ld3(p.Uold, c3, TS);
        c.dOm = p.first ? c.Om : ld3(p.Omold, c3, TS);
    }
    return c;
}
template <int SCHEME>
DEV void addLoadsAndInertia(const BeamParams& p, const int b, const int i, const double dZ, const double ARho, const V3 CI,
                            const CellIn& cell, double (&Dc)[6][6], double (&src)[6])
{
    const double dt = p.dt, gamma = p.gamma;
    // ---- loads: Evolve.C:193-209,274-279
    const size_t c3 = tileRow((size_t)(b >> 5), p.nmax, i, 3, b & 31);
    const double Lcell = p.Lc ? p.Lc[tileRow((size_t)(b >> 5), p.nmax, i, 1, b & 31)] : dZ;
    if (p.q) subV(src, 0, Lcell * ld3(p.q, c3, TS));
    if (p.weight) subV(src, 0, Lcell * ld3(p.weight, b, p.Bpad)); // null when no beam has a self-weight
    if (p.m) subV(src, 3, Lcell * ld3(p.m, c3, TS));
    if (p.pfSrc) { // point forces and momentum contributions, formed by their pre-kernels
        const size_t c6 = tileRow((size_t)(b >> 5), p.nmax, i, 6, b & 31);
#pragma unroll
        for (int r = 0; r < 6; r++) src[r] -= p.pfSrc[c6 + (size_t)r * TS];
    }
    if (SCHEME == BFK_NEWMARK) { // inertia: Evolve.C:320-536 (+ predictor :167-186 on the first iteration)
        const M3 Lm = quatToMat(cell.ql);
        V3 Ac = cell.Ac, dOm = cell.dOm, Om;
        const V3 OmS = cell.Om; // Omega*
        if (p.first) {
            const V3 Ao = Ac, dOo = dOm;
            const V3 Uo = cell.U + p.gdtPrev * Ao, Oo = OmS + p.gdtPrev * dOo; // end of the previous step
            const double k1 = p.nk1, k2 = p.nk2;
            Ac = v3(-k1 * Uo.x - k2 * Ao.x, -k1 * Uo.y - k2 * Ao.y, -k1 * Uo.z - k2 * Ao.z);
            dOm = v3(-k1 * Oo.x - k2 * dOo.x, -k1 * Oo.y - k2 * dOo.y, -k1 * Oo.z - k2 * dOo.z);
            Om = Oo + dt * ((1.0 - gamma) * dOo + gamma * dOm);
        } else
            Om = OmS + p.gdt * dOm;
        // QRho = ARho L Accl, MRho = L Lambda (CI dotOmega + Omega x CI Omega) :369-419;  QRhoCoeff, MRhoCoeff :459-535:
        //   MRhoCoeff/L = S(a) + g1 (S(Lambda CI Omega) - Lambda S(Omega) CI Lambda.T()) - g2 Lambda CI Lambda.T()
        //               = S(a + g1 Lambda CI Omega) + Lambda (-(g1 S(Omega) + g2 I) CI Lambda.T()),  g1 = gamma/(beta dt), g2 = 1/(beta dt^2)
        const V3 CIOm = cmul(CI, Om), CIdO = cmul(CI, dOm);
        const V3 w = v3(fma(-Om.z, CIOm.y, fma(Om.y, CIOm.z, CIdO.x)), fma(-Om.x, CIOm.z, fma(Om.z, CIOm.x, CIdO.y)),
                        fma(-Om.y, CIOm.x, fma(Om.x, CIOm.y, CIdO.z)));
        const V3 a = mul(Lm, w);
        const double mL = ARho * Lcell;
        src[0] = fma(mL, Ac.x, src[0]); src[1] = fma(mL, Ac.y, src[1]); src[2] = fma(mL, Ac.z, src[2]);
        src[3] = fma(Lcell, a.x, src[3]); src[4] = fma(Lcell, a.y, src[4]); src[5] = fma(Lcell, a.z, src[5]);
        const double QRhoCoeff = -mL * p.nAcc;
        Dc[0][0] += QRhoCoeff; Dc[1][1] += QRhoCoeff; Dc[2][2] += QRhoCoeff;
        const M3 Bt = diagMulT(CI, Lm);             // CI Lambda.T()
        const M3 SB = mulSpinL(Om, Bt);             // S(Omega) CI Lambda.T()
        M3 N;
#pragma unroll
        for (int k = 0; k < 9; k++) N.a[k] = fma(-p.nVel, SB.a[k], -p.nAcc * Bt.a[k]);
        const M3 LN = mul(Lm, N);
        const V3 sv = a + p.nVel * mul(Lm, CIOm);
        Dc[3][3] = fma(Lcell, LN.a[0], Dc[3][3]); Dc[3][4] = fma(Lcell, LN.a[1] - sv.z, Dc[3][4]); Dc[3][5] = fma(Lcell, LN.a[2] + sv.y, Dc[3][5]);
        Dc[4][3] = fma(Lcell, LN.a[3] + sv.z, Dc[4][3]); Dc[4][4] = fma(Lcell, LN.a[4], Dc[4][4]); Dc[4][5] = fma(Lcell, LN.a[5] - sv.x, Dc[4][5]);
        Dc[5][3] = fma(Lcell, LN.a[6] - sv.y, Dc[5][3]); Dc[5][4] = fma(Lcell, LN.a[7] + sv.x, Dc[5][4]); Dc[5][5] = fma(Lcell, LN.a[8], Dc[5][5]);
    }
    if (SCHEME == BFK_EULER) { // Evolve.C:385-400, 490-513 with fvc::ddt(X) = (X - X.oldTime())/deltaT
        const double idt = 1.0 / dt;
        const M3 Lm = quatToMat(cell.ql);
        const V3 U = cell.U, Om = cell.Om, Uold = cell.Ac, Omold = cell.dOm; // see CellIn
        const V3 CIOm = cmul(CI, Om);
        const V3 gyro = cross(Om, CIOm);
        const V3 a = mul(Lm, cmul(CI, idt * (Om - Omold)) + gyro);
        const double mL = ARho * Lcell;
        addV(src, 0, (mL * idt) * (U - Uold));
        addV(src, 3, Lcell * a);
        const double QRhoCoeff = -mL * idt * idt;
        Dc[0][0] += QRhoCoeff; Dc[1][1] += QRhoCoeff; Dc[2][2] += QRhoCoeff;
        // MRhoCoeff/L = -Lam CI Lam.T/dt^2 - S(Lam CI Om.old)/dt - S(Lam S(Om) CI Om) + Lam S(Om) CI Lam.T/dt
        // (the +S(Lam CI Om)/dt and -S(Lam CI Om)/dt of :500,:509 cancel)
        const M3 Bt = diagMulT(CI, Lm);
        const M3 SB = mulSpinL(Om, Bt);
        M3 N;
#pragma unroll
        for (int k = 0; k < 9; k++) N.a[k] = idt * SB.a[k] - (idt * idt) * Bt.a[k];
        const M3 MR = mul(Lm, N) - spin(mul(Lm, idt * cmul(CI, Omold) + gyro));
        addBlk(Dc, 3, 3, Lcell, MR);
    }
}

// Assembly of cell i, face part (first third of a forward step).  Takes the rows of face i+1 and W_{i+1} (from the
// prefetch buffer, already in registers), adds the own part of d_i to the carry Dc, starts source_i, and parks the face
// tensors in the stash.  LAST = the cell whose right face is the right patch: it closes the patch instead of an
// internal face.  The loads and inertia of the cell follow in addLoadsAndInertia once its Newmark rows have arrived.
// MIR: the cell belongs to the mirrored half of a split beam (forwardSweepSplit): the sweep walks right to left, "the
// next cell" is cell i-1 and the face between them is seen from its neighbour side.  The neighbour-side blocks are the
// owner-side blocks with Qt and M negated (d_nei = [[-P, -Qt], [-Qt.T(), S3 - S2 - S1]], l = [[P, -Qt], [Qt.T(), S1 + S3 - S2]],
// source parts (+eQ ; eM - eMQ)), i.e. the same code with th, hQ, hM, eQ, eM negated; t keeps the face's own orientation.
template <bool LAST, bool MIR = false>
DEV void fwdFacePart(const BeamParams& p, const int b, double* __restrict__ sm, const double dcI, const double h,
                     const FaceIn& in, double (&Dc)[6][6], double (&src)[6], V3& Wi)
{
    // Dc is an OUTPUT: each 3x3 block of the carry is fetched from the stash (R_DC) where its first
    // contribution is added
#pragma unroll
    for (int r = 0; r < 6; r++) src[r] = smLd(sm, R_SC + r);
    if (!LAST) {
        // ---- coefficients of face i+1 (Evolve.C:673-753) and own part of d_i (Matrix.C:84-470, w = 0.5, a = 1/deltaf).
        // The scalings of the reference (a = deltaCoeffs, 0.5/deltaCoeffs, the 0.5 interpolation weights) are folded
        // into the vectors: with h = 0.5/deltaCoeffs, th = h t, hQ = 0.5 Q, hM = 0.5 M and P = a CQW,
        //   0.5 CQTheta = P S(th) - S(hQ)            a CMQW = -(0.5 CQTheta).T()        explicitMQ = th x explicitQ
        //   0.5 CMQTheta = -(Gh.T() S(th) + S(th) S(hQ)),  Gh = P S(th),  S(a) S(b) = b a.T() - (a.b) I
        const V3 Wn = in.Wn;
        const V3 dR0 = smLd3(sm, R_DR0);
        const double sdc = MIR ? -dcI : dcI;
        const V3 t = v3(dR0.x + sdc * (Wn.x - Wi.x), dR0.y + sdc * (Wn.y - Wi.y), dR0.z + sdc * (Wn.z - Wi.z)); // dR0Ds + snGrad(W)
        Wi = Wn;
        const V3 th = (MIR ? -h : h) * t;
        const M3 Lm = quatToMat(in.qf);                                               // Lambdaf & refLambdaf :678
        const V3 hQ = (MIR ? -0.5 : 0.5) * in.Q;
        V3 eQ;
        M3 Gh;
        {
            const V3 CQ = smLd3(sm, R_CQ);
            const Sym3 Pm = conjDiagSym(Lm, dcI * CQ);                                // a*CQW :680
            // refLambdaf.T() & dR0Ds = e_x: bf_create orthonormalises refLambdaf and dR0Ds is its first column
            eQ = mul(Lm, cmul(CQ, mulTSub(Lm, t, v3(1.0, 0.0, 0.0))));                // :682 with Gamma of :881-887
            if (MIR) eQ = -eQ;
            Gh = mulSpinR(Pm, th);
            M3 Qt = Gh;                                                               // 0.5*CQTheta :686-693
            Qt.a[1] += hQ.z; Qt.a[2] -= hQ.y; Qt.a[3] -= hQ.z; Qt.a[5] += hQ.x; Qt.a[6] += hQ.y; Qt.a[7] -= hQ.x;
            smStS(sm, R_P, Pm);
            smSt9(sm, R_QT, Qt);
            carryBlk(Dc, sm, 0, 0, -1.0, full(Pm));
            carryBlk(Dc, sm, 0, 3, 1.0, Qt);
            carryBlk(Dc, sm, 3, 0, 1.0, transpose(Qt));                               // -a*CMQW :706-715
        }
        const V3 eMQ = cross(th, eQ);                                                 // :737-741
        {
            const V3 CM = smLd3(sm, R_CM);
            const Sym3 S1 = conjDiagSym(Lm, dcI * CM);                                // a*CMTheta :699
            V3 eM = mul(Lm, cmul(CM, in.K));                                          // :684
            if (MIR) eM = -eM;
            M3 S3;                                                                    // 0.5*CMQTheta :718-735
            {
                const double d = dot(th, hQ);
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const V3 y = cross(th, v3(Gh.a[r], Gh.a[3 + r], Gh.a[6 + r])); // -(column r of Gh) x th
                    const double q = r == 0 ? hQ.x : r == 1 ? hQ.y : hQ.z;
                    S3.a[3 * r] = fma(-q, th.x, y.x); S3.a[3 * r + 1] = fma(-q, th.y, y.y); S3.a[3 * r + 2] = fma(-q, th.z, y.z);
                    S3.a[4 * r] += d;
                }
            }
            const V3 hM = (MIR ? -0.5 : 0.5) * in.M;                                  // 0.5*CMTheta2 = -S(hM) :703
            smStS(sm, R_S1, S1);
            smSt9(sm, R_S3, S3);
            smSt3(sm, R_MF, hM);
            M3 Eo = S3 - full(S1);                                                    // -a*CMTheta + 0.5*CMQTheta + 0.5*CMTheta2
            Eo.a[1] += hM.z; Eo.a[2] -= hM.y; Eo.a[3] -= hM.z; Eo.a[5] += hM.x; Eo.a[6] += hM.y; Eo.a[7] -= hM.x;
            carryBlk(Dc, sm, 3, 3, 1.0, Eo);
            subV(src, 0, eQ);
            subV(src, 3, eM + eMQ);
            smSt3(sm, R_SC, eQ);
            smSt3(sm, R_SC + 3, eM - eMQ);
        }
    } else { // right patch
        double D36[36], s6[6];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            s6[r] = src[r];
#pragma unroll
            for (int c = 0; c < 6; c++) D36[6 * r + c] = smLd(sm, R_DC + 6 * r + c);
        }
        patchClosure(p, 1, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            src[r] = s6[r];
#pragma unroll
            for (int c = 0; c < 6; c++) Dc[r][c] = D36[6 * r + c];
        }
    }
}

// Fallback of the elimination step when a 3x3 pivot block of D'_i is ill-conditioned (smallDet, beam_math.cuh): the whole
// 6x6 block by LU with partial pivoting, the way the reference's SparseLU treats it (EigenOF.C:292-297).  The right-hand
// sides are rebuilt from the stash exactly as in fwdEliminate.  X42 = D'^{-1} [u | source - lb], row-major [6][7].
__device__ __noinline__ void fwdSolvePivoted(const double* __restrict__ sm, const double* __restrict__ D36,
                                             const double* __restrict__ src6, const int last, double* __restrict__ X42)
{
    double A[6][6], Bq[6][7];
    const int pi[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}};
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 6; c++) A[r][c] = D36[6 * r + c];
    const V3 hM = smLd3(sm, R_MF);
#pragma unroll
    for (int c = 0; c < 7; c++) {
        V3 r1 = v3(0, 0, 0), r2 = v3(0, 0, 0);
        if (c < 3) {
            if (!last) {
                r1 = v3(smLd(sm, R_P + pi[c][0]), smLd(sm, R_P + pi[c][1]), smLd(sm, R_P + pi[c][2]));
                r2 = -smLd3(sm, R_QT + 3 * c);
            }
        } else if (c < 6) {
            if (!last) {
                const int k = c - 3;
                const V3 s2k = k == 0 ? v3(0.0, -hM.z, hM.y) : k == 1 ? v3(hM.z, 0.0, -hM.x) : v3(-hM.y, hM.x, 0.0);
                r1 = v3(smLd(sm, R_QT + k), smLd(sm, R_QT + 3 + k), smLd(sm, R_QT + 6 + k));
                r2 = v3(smLd(sm, R_S1 + pi[k][0]) + smLd(sm, R_S3 + k) + s2k.x, smLd(sm, R_S1 + pi[k][1]) + smLd(sm, R_S3 + 3 + k) + s2k.y,
                        smLd(sm, R_S1 + pi[k][2]) + smLd(sm, R_S3 + 6 + k) + s2k.z);
            }
        } else {
            r1 = v3(src6[0] + smLd(sm, R_LB), src6[1] + smLd(sm, R_LB + 1), src6[2] + smLd(sm, R_LB + 2));
            r2 = v3(src6[3] + smLd(sm, R_LB + 3), src6[4] + smLd(sm, R_LB + 4), src6[5] + smLd(sm, R_LB + 5));
        }
        Bq[0][c] = r1.x; Bq[1][c] = r1.y; Bq[2][c] = r1.z; Bq[3][c] = r2.x; Bq[4][c] = r2.y; Bq[5][c] = r2.z;
    }
    solve6<7>(A, Bq);
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
        for (int c = 0; c < 7; c++) X42[7 * r + c] = Bq[r][c];
}

// Elimination step of cell i (second half): D'_i = Dc is factored by 2x2 block elimination over
// 3x3 blocks, X = D'^{-1} [u_i | source_i - l_{i-1} bPrime_{i-1}] is formed column by column,
// stored (uPrime, bPrime), and folded straight into the carry of cell i+1:
//   l = [[P, -Qt], [Qt.T(), S1 + Sm]],   d_nei = [[-P, -Qt], [-Qt.T(), Sm - S1]],   Sm = S3 + 0.5 spin(M)
//   Dc <- d_nei - l & uPrime ;  lb <- l & bPrime
template <bool LAST>
DEV void fwdEliminate(const BeamParams& p, const int b, double* __restrict__ sm, const int i, const double (&Dc)[6][6],
                      const double (&src)[6])
{
    M3 Ai, AiB, Si, Cm;
    bool illConditioned;
    {
        M3 A, Bm, E;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                A.a[3 * r + c] = Dc[r][c];
                Bm.a[3 * r + c] = Dc[r][3 + c];
                Cm.a[3 * r + c] = Dc[3 + r][c];
                E.a[3 * r + c] = Dc[3 + r][3 + c];
            }
        double detA, detS, mA, mS;
        Ai = invDet(A, detA, mA);
        AiB = mul(Ai, Bm);
        Si = invDet(subMulM(E, Cm, AiB), detS, mS);
        illConditioned = smallDet(detA, mA) || smallDet(detS, mS);
    }
    // ---- X = D'^{-1} [u | source - lb], column by column; u comes back from the stash.  R_LB holds -lb.
    double X1[3][7], X2[3][7];
    const int pi[3][3] = {{0, 1, 2}, {1, 3, 4}, {2, 4, 5}}; // symmetric storage: xx xy xz yy yz zz
    if (illConditioned) { // a 3x3 pivot block is (near-)singular: pivoted LU of the whole 6x6 block, out of line
        double D36[36], s6[6], X42[42];
#pragma unroll
        for (int r = 0; r < 6; r++) {
            s6[r] = src[r];
#pragma unroll
            for (int c = 0; c < 6; c++) D36[6 * r + c] = Dc[r][c];
        }
        fwdSolvePivoted(sm, D36, s6, LAST, X42);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 7; c++) { X1[r][c] = X42[7 * r + c]; X2[r][c] = X42[7 * (r + 3) + c]; }
    } else {
#pragma unroll
    for (int c = 0; c < 7; c++) {
        if (c < 6 && LAST) continue;
        V3 r1, r2;
        if (c < 3) { // u[:, c] = (P[:, c] ; -Qt[c, :])
            r1 = v3(smLd(sm, R_P + pi[c][0]), smLd(sm, R_P + pi[c][1]), smLd(sm, R_P + pi[c][2]));
            r2 = -smLd3(sm, R_QT + 3 * c);
        } else if (c < 6) { // u[:, c] = (Qt[:, k] ; (S1 + S3 + S2)[:, k]),  S2 = -spin(hM)
            const int k = c - 3;
            const V3 hM = smLd3(sm, R_MF);
            const V3 s2k = k == 0 ? v3(0.0, -hM.z, hM.y) : k == 1 ? v3(hM.z, 0.0, -hM.x) : v3(-hM.y, hM.x, 0.0);
            r1 = v3(smLd(sm, R_QT + k), smLd(sm, R_QT + 3 + k), smLd(sm, R_QT + 6 + k));
            r2 = v3(smLd(sm, R_S1 + pi[k][0]) + smLd(sm, R_S3 + k) + s2k.x, smLd(sm, R_S1 + pi[k][1]) + smLd(sm, R_S3 + 3 + k) + s2k.y,
                    smLd(sm, R_S1 + pi[k][2]) + smLd(sm, R_S3 + 6 + k) + s2k.z);
        } else { // source_i - l_{i-1} & bPrime_{i-1}
            r1 = v3(src[0] + smLd(sm, R_LB), src[1] + smLd(sm, R_LB + 1), src[2] + smLd(sm, R_LB + 2));
            r2 = v3(src[3] + smLd(sm, R_LB + 3), src[4] + smLd(sm, R_LB + 4), src[5] + smLd(sm, R_LB + 5));
        }
        const V3 y1 = mul(Ai, r1);
        const V3 x2 = mul(Si, subMul(r2, Cm, y1));
        const V3 x1 = subMul(y1, AiB, x2);
        X1[0][c] = x1.x; X1[1][c] = x1.y; X1[2][c] = x1.z;
        X2[0][c] = x2.x; X2[1][c] = x2.y; X2[2][c] = x2.z;
    }
    }
    // ---- store uPrime (36) and bPrime (6)
    const size_t u0 = tileRow((size_t)(b >> 5), p.nmax, i, 36, b & 31), b0 = tileRow((size_t)(b >> 5), p.nmax, i, 6, b & 31);
    if (!LAST) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                stS(p.uP + u0 + (size_t)(6 * r + c) * TS, X1[r][c]);
                stS(p.uP + u0 + (size_t)(6 * (r + 3) + c) * TS, X2[r][c]);
            }
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
        stS(p.bP + b0 + (size_t)r * TS, X1[r][6]);
        stS(p.bP + b0 + (size_t)(r + 3) * TS, X2[r][6]);
    }
    if (!LAST) {
        // ---- carry for cell i+1, one row of l at a time, each entry one fused chain that starts from d_nei:
        //      Dc[r][c] = d_nei[r][c] - sum_k l[r][k] X[k][c];   -lb[r] = 0 - sum_k l[r][k] X[k][6]
        {
            const M3 Pm = full(smLdS(sm, R_P)), Qt = smLd9(sm, R_QT);
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int c = 0; c < 7; c++) {
                    double acc = c < 3 ? -Pm.a[3 * r + c] : c < 6 ? -Qt.a[3 * r + c - 3] : 0.0;
#pragma unroll
                    for (int k = 0; k < 3; k++) acc = fma(-Pm.a[3 * r + k], X1[k][c], acc);
#pragma unroll
                    for (int k = 0; k < 3; k++) acc = fma(Qt.a[3 * r + k], X2[k][c], acc);
                    if (c < 6) smSt(sm, R_DC + 6 * r + c, acc);
                    else smSt(sm, R_LB + r, acc);
                }
            }
        }
        {
            const M3 Qt = smLd9(sm, R_QT), S1 = full(smLdS(sm, R_S1));
            const V3 hM = smLd3(sm, R_MF);
            M3 Sm = smLd9(sm, R_S3); // S3 - S2 = S3 + spin(hM)
            Sm.a[1] -= hM.z; Sm.a[2] += hM.y; Sm.a[3] += hM.z;
            Sm.a[5] -= hM.x; Sm.a[6] -= hM.y; Sm.a[7] += hM.x;
            const M3 Sl = S1 + Sm, Sd = Sm - S1;
#pragma unroll
            for (int r = 0; r < 3; r++) {
#pragma unroll
                for (int c = 0; c < 7; c++) {
                    double acc = c < 3 ? -Qt.a[3 * c + r] : c < 6 ? Sd.a[3 * r + c - 3] : 0.0;
#pragma unroll
                    for (int k = 0; k < 3; k++) acc = fma(-Sl.a[3 * r + k], X2[k][c], acc);
#pragma unroll
                    for (int k = 0; k < 3; k++) acc = fma(-Qt.a[3 * k + r], X1[k][c], acc);
                    if (c < 6) smSt(sm, R_DC + 6 * (3 + r) + c, acc);
                    else smSt(sm, R_LB + 3 + r, acc);
                }
            }
        }
    }
}

// The last cell of a beam (right patch instead of an internal face): once per beam, kept out of line so that
// its patch code and stack arrays stay out of the register allocation of the steady-state loop.
template <int SCHEME>
__device__ __noinline__ double fwdLastCell(const BeamParams& p, const int b, double* __restrict__ sm, const int i,
                                           const double dZ, const double ARho, V3 Wi, const CellIn* cell)
{
    double Dc[6][6], src[6], res = 0.0;
    FaceIn none;
    none.Wn = none.K = none.Q = none.M = v3(0, 0, 0); none.qf = q4(1, 0, 0, 0);
    fwdFacePart<true>(p, b, sm, 1.0 / dZ, 0.5 * dZ, none, Dc, src, Wi);
    addLoadsAndInertia<SCHEME>(p, b, i, dZ, ARho, smLd3(sm, R_CI), *cell, Dc, src);
#pragma unroll
    for (int r = 0; r < 6; r++) res += src[r] * src[r];
    fwdEliminate<true>(p, b, sm, i, Dc, src);
    return res;
}

// One forward step for a cell that is not interior in every lane of the warp (see forwardSweep): same protocol on the
// two barriers, per-lane branches around the per-lane parts.  Out of line: it runs once per beam in a uniform bundle.
template <int SCHEME>
__device__ __noinline__ void fwdCellGeneric(const BeamParams& p, const int b, double* __restrict__ warpSm, uint64_t* bar,
                                            const int i, const int nW, uint32_t* phase, V3* WiIO, double* resIO)
{
    const size_t T = (size_t)(b >> 5);
    double* sm = warpSm + (threadIdx.x & 31);
    V3 Wi = *WiIO;
    double res = 0.0;
    mbarWait(bar, *phase & 1u);
    const FaceIn in = loadFaceIn(sm);
    if (SCHEME == BFK_STEADY) fenceProxyAsync();
    __syncwarp();
    if (SCHEME != BFK_STEADY) fwdPrefetchCell<SCHEME>(p, T, sm, i);
    else if (i + 1 < nW) fwdPrefetchFace<SCHEME>(p, T, warpSm, bar, i + 1);
    const int nb = (int)smLd(sm, R_N); // cells of this lane's beam
    const bool interior = i < nb - 1;  // a cell with an internal face on its right
    double src[6], Dc[6][6];
    if (interior) {
        const double dz = smLd(sm, R_DZ);
        fwdFacePart<false>(p, b, sm, rcpFast(dz), 0.5 * dz, in, Dc, src, Wi);
    }
    CellIn cell;
    if (SCHEME != BFK_STEADY) {
        cpAsyncWaitAll();
        cell = loadCellInSm<SCHEME>(sm, p.first != 0);
        fenceProxyAsync();
        __syncwarp();
        if (i + 1 < nW) fwdPrefetchFace<SCHEME>(p, T, warpSm, bar, i + 1);
    } else
        cell = loadCellIn<SCHEME>(p, b, i);
    *phase += 1;
    if (interior) {
        addLoadsAndInertia<SCHEME>(p, b, i, smLd(sm, R_DZ), smLd(sm, R_ARHO), smLd3(sm, R_CI), cell, Dc, src);
#pragma unroll
        for (int r = 0; r < 6; r++) res += src[r] * src[r];
        fwdEliminate<false>(p, b, sm, i, Dc, src);
    } else if (i == nb - 1)
        res += fwdLastCell<SCHEME>(p, b, sm, i, smLd(sm, R_DZ), smLd(sm, R_ARHO), Wi, &cell);
    *WiIO = Wi;
    *resIO += res;
}

// Forward sweep of the 32 beams of one warp.  sm = the warp's slab + lane; bar = the warp's mbarrier for the face rows.
// phase = number of cells this warp has processed so far: the barrier completes once per cell, its parity selects the wait.
template <int SCHEME>
DEV double forwardSweep(const BeamParams& p, const int b, double* __restrict__ warpSm, uint64_t* bar, uint32_t& phase)
{
    const size_t s = p.Bpad;
    const int lane = threadIdx.x & 31;
    const size_t T = (size_t)(b >> 5);
    double* sm = warpSm + lane;
    double res = 0.0;
    const bool live = b < p.B;
    const int n = live ? p.nSeg[b] : 0;
    const int nW = __reduce_max_sync(0xffffffffu, n); // cells of the longest beam of the warp
    __syncwarp(); // (persistent blocks) every lane has left the previous tile's buffer
    if (nW > 0) fwdPrefetchFace<SCHEME>(p, T, warpSm, bar, 0);

    {
        const M3 refL = ld9(p.refL, b, s);
        const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
        smSt3(sm, R_DR0, dR0);
        smSt(sm, R_DZ, p.dZ[b]);
        smSt(sm, R_ARHO, SCHEME != BFK_STEADY ? p.ARho[b] : 0.0);
        smSt(sm, R_N, (double)n);
        smSt3(sm, R_CQ, ld3(p.CQ, b, s));
        smSt3(sm, R_CM, ld3(p.CM, b, s));
        smSt3(sm, R_CI, SCHEME != BFK_STEADY ? ld3(p.CI, b, s) : v3(0, 0, 0));
#pragma unroll
        for (int r = 0; r < 6; r++) smSt(sm, R_LB + r, 0.0);
    }
    // loop-carried in registers: lb = l_{i-1} & bPrime_{i-1};  Wi = W of cell i.  The carry
    // (nei part of d_i) - l_{i-1} & uPrime_{i-1} and the nei part of source_i travel through the stash.
    V3 Wi = v3(0, 0, 0);
    if (live) {
        Wi = ld3(p.W, tileRow(T, p.nmax, 0, 3, lane), TS);
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        patchClosure(p, 0, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            smSt(sm, R_SC + r, s6[r]);
#pragma unroll
            for (int c = 0; c < 6; c++) smSt(sm, R_DC + 6 * r + c, D36[6 * r + c]);
        }
    }
    // Cells 0 .. nFast-1 have an internal face on their right in EVERY lane of the warp: straight-line code, the
    // warp-collective steps (barrier waits, buffer hand-over) sit between the per-lane parts without any divergence.
    // The remaining cells (the last cell of each beam; all cells past the shortest beam of a ragged warp) take the
    // generic out-of-line step.
    const int nFast = __reduce_min_sync(0xffffffffu, n) - 1;
    int i = 0;
    for (; i < nFast; i++) {
        mbarWait(bar, phase & 1u);
        const FaceIn in = loadFaceIn(sm);
        if (SCHEME == BFK_STEADY) fenceProxyAsync();
        __syncwarp(); // every lane holds its buffer rows in registers: the buffer is free for the next fill
        if (SCHEME != BFK_STEADY) fwdPrefetchCell<SCHEME>(p, T, sm, i);
        else fwdPrefetchFace<SCHEME>(p, T, warpSm, bar, i + 1);
        double src[6], Dc[6][6];
        const double dz = smLd(sm, R_DZ);
        fwdFacePart<false>(p, b, sm, rcpFast(dz), 0.5 * dz, in, Dc, src, Wi);
        CellIn cell;
        if (SCHEME != BFK_STEADY) { // the cell rows of cell i have arrived; hand the buffer to face i+2 / cell i+1
            cpAsyncWaitAll();
            cell = loadCellInSm<SCHEME>(sm, p.first != 0);
            fenceProxyAsync();
            __syncwarp();
            fwdPrefetchFace<SCHEME>(p, T, warpSm, bar, i + 1);
        } else
            cell = loadCellIn<SCHEME>(p, b, i);
        phase++;
        addLoadsAndInertia<SCHEME>(p, b, i, dz, smLd(sm, R_ARHO), smLd3(sm, R_CI), cell, Dc, src);
#pragma unroll
        for (int r = 0; r < 6; r++) res += src[r] * src[r]; // ||source||^2 (EigenOF.C:253-282, x0 = 0)
        fwdEliminate<false>(p, b, sm, i, Dc, src);
    }
    for (; i < nW; i++) fwdCellGeneric<SCHEME>(p, b, warpSm, bar, i, nW, &phase, &Wi, &res);
    return res;
}

// ---------------------------------------------------------------------------
// backward: back-substitution + BC evaluate + updateSolutionVariables + norms
// ---------------------------------------------------------------------------
struct FaceState {
    Q4 Lf;
    V3 K, Q, M;
};

// Update of one face (Evolve.C:804-836, 881-913).  (A) = owner / face cell, (B) = neighbour /
// patch value.  f holds the PRE-solve coefficients; tnew = dR0Ds + snGrad(W_new).
DEV FaceState updateFace(const FaceCoef& f, Q4 qOld, const M3& LmOld, V3 Kold,
                         V3 DWa, V3 DWb, V3 DTa, V3 DTb, double dc, bool boundary)
{
    FaceState o;
    const V3 DThf = boundary ? DTb : v3(0.5 * DTa.x + 0.5 * DTb.x, 0.5 * DTa.y + 0.5 * DTb.y, 0.5 * DTa.z + 0.5 * DTb.z);
    const V3 sgT = dc * (DTb - DTa), sgW = dc * (DWb - DWa);
    const HalfAngle h = halfAngle(DThf);
    // K += (refLambdaf.T() & Lambdaf.T()) & (DT.T() & snGrad(DTheta))   :825-830 (old Lambdaf)
    o.K = Kold + mulT(LmOld, tangentT(DThf, sgT, h));
    // Lambdaf = R(DThetaf) & Lambdaf :833-836; the same left factor updates the stored Lambdaf & refLambdaf
    o.Lf = quatNormalize(quatMul(quatFromRotVec(DThf, h), qOld));
    // Gamma (:881-887) is a function of the new W and Lambdaf: recomputed by its consumers, not stored
    // Q = explicitQ + CQW & snGrad(DW) + CQTheta & interpolate(DTheta)  :896-903
    o.Q = f.eQ + (mul(f.CQW, sgW) + mul(f.CQTheta, DThf));
    // M = explicitM + CMTheta & snGrad(DTheta) + CMTheta2 & interpolate(DTheta) :906-913
    o.M = f.eM + (mul(f.CMTheta, sgT) - cross(f.Mold, DThf));
    return o;
}

// correctBoundaryConditions() for ONE end patch followed by the update of its boundary face.
// Theta BC first, then W BC (Evolve.C:764 then :771), so the W evaluate sees the new DTheta_b.
//   momentBeamRotationNR::evaluate        PF/momentBeamRotationNR/...C:183-222
//   forceBeamDisplacementNR::evaluate     PF/forceBeamDisplacementNR/...C:218-262 (also followerForce)
//   linearSpringForce...::evaluate        PF/linearSpringForceBeamDisplacementNR/...C:219-300
//   fixedValue / fixedRotation / fixedDisplacement: value from updateCoeffs(); DW_b, DTheta_b are the
//   pWCorr / pThetaCorr written by assembleBoundaryConditions (BC.C:359-361, 478-482).
// dR0s is dR0Ds of this patch (sign-flipped on the left patch).
__device__ __noinline__ FaceState patchUpdate(const BeamParams& p, int end, int b, Q4 qOld, V3 CQ, V3 CM,
                          V3 Kold, V3 Q, V3 M, V3 WoldC, V3 DWc, V3 DTc, double dcB, double delta)
{
    const size_t s = p.Bpad;
    const M3 refL = ld9(p.refL, b, s);
    const V3 dR0s = end ? v3(refL.a[0], refL.a[3], refL.a[6]) : v3(-refL.a[0], -refL.a[3], -refL.a[6]);
    const size_t i3 = (size_t)end * 3 * s + b, i9 = (size_t)end * 9 * s + b;
    const M3 LmOld = quatToMat(qOld);       // Lambdaf_b & refLambdaf
    const M3 LfOld = mulBT(LmOld, refL);    // Lambdaf_b
    const EndBC e = loadEnd(p, end, b, LfOld);
    const V3 WbOld = e.wType == BFK_W_FIXED ? e.wPresc : e.Wb;
    const V3 t = dR0s + dcB * (WbOld - WoldC);
    const V3 Gam = gammaOf(LmOld, mulT(refL, dR0s), dR0s + dcB * (e.Wb - WoldC));
    const FaceCoef f = faceCoefficients(LmOld, CQ, CM, Gam, Kold, Q, M, t, dcB, true);
    const double ipd = 1.0 / delta;
    const M3 LambOld = ld9(p.Lamb, i9, s);
    V3 DTb, ThbNew;
    if (e.thType == BFK_TH_MOMENT) {
        const M3 A = ipd * f.CMTheta;
        const M3 invA = inv(A - spin(f.Mold)); // CMTheta/delta + CMTheta2
        DTb = mul(invA, e.moment - f.eM) + mul(mul(invA, A), DTc);
        ThbNew = e.Thb + mulT(LambOld, DTb);
    } else {
        DTb = mul(LfOld, e.thPresc - e.Thb);
        ThbNew = e.thPresc;
    }
    V3 DWb, WbNew;
    if (e.wType == BFK_W_FIXED) {
        DWb = e.wPresc - e.Wb;
        WbNew = e.wPresc;
    } else {
        M3 A = ipd * f.CQW;
        V3 force = e.force;
        if (e.wType == BFK_W_SPRING) {
            const V3 sd = ld3(p.springDir, i3, s);
            const double k_ = p.springK[(size_t)end * s + b];
            const double fm = k_ * dot(e.Wb, sd); // from the boundary W of the previous iterate
            force = ld3(p.offset, i3, s) - fm * sd;
            st3(p.force, i3, s, force);
            M3 Am = A;
            Am.a[0] += k_ * (sd.x * sd.x); Am.a[1] += k_ * (sd.x * sd.y); Am.a[2] += k_ * (sd.x * sd.z);
            Am.a[3] += k_ * (sd.y * sd.x); Am.a[4] += k_ * (sd.y * sd.y); Am.a[5] += k_ * (sd.y * sd.z);
            Am.a[6] += k_ * (sd.z * sd.x); Am.a[7] += k_ * (sd.z * sd.y); Am.a[8] += k_ * (sd.z * sd.z);
            const M3 invA = inv(Am);
            DWb = mul(invA, force - f.eQ) - mul(invA, mul(f.CQTheta, DTb)) + mul(invA, mul(A, DWc));
        } else {
            const M3 invA = inv(A);
            DWb = mul(invA, force - f.eQ) - mul(invA, mul(f.CQTheta, DTb)) + DWc; // CQDTheta == 0
        }
        WbNew = e.Wb + DWb;
    }
    st3(p.Wb, i3, s, WbNew);
    st3(p.Thb, i3, s, ThbNew);
    st3(p.DWb, i3, s, DWb);
    st3(p.DThb, i3, s, DTb);
    st9(p.Lamb, i9, s, mul(rotationMatrix(DTb), LambOld)); // Lambda_ = DLambda & Lambda_ on the patch
    return updateFace(f, qOld, LmOld, Kold, DWc, DWb, DTc, DTb, dcB, true);
}

// The rows of a cell that its update reads on every iteration (Newmark: + Accl, dotOmega).
struct CellOld {
    V3 W, Th, Ac, dOm;
    Q4 q;
};
template <int SCHEME>
DEV CellOld loadCellOld(const BeamParams& p, const size_t T, const int lane, const int i)
{
    CellOld o;
    const size_t c3 = tileRow(T, p.nmax, i, 3, lane), c4 = tileRow(T, p.nmax, i, 4, lane);
    o.W = ld3(p.W, c3, TS);
    o.q = ldq(p.Lam, c4, TS);
    o.Th = ld3(p.Th, c3, TS);
    o.Ac = o.dOm = v3(0, 0, 0);
    if (SCHEME == BFK_NEWMARK) { o.Ac = ld3(p.Ac, c3, TS); o.dOm = ld3(p.dOm, c3, TS); }
    return o;
}
template <int SCHEME>
DEV void updateCellFrom(const BeamParams& p, const size_t T, const int lane, const int i, const V3 DW, const V3 DT, const CellOld& old,
                        double& x2)
{
    const size_t s = TS;
    const int Rc = p.nmax;
    const double dt = p.dt, gamma = p.gamma;
    // ---- cell i: Evolve.C:759-802, 839-863
    const size_t c3 = tileRow(T, Rc, i, 3, lane), c4 = tileRow(T, Rc, i, 4, lane);
    const V3 Wold = old.W;
    const Q4 qOld = old.q;
    // Lambda.T() & DTheta.  R(DTheta) leaves its own axis unchanged, so this vector is the same for
    // the Lambda before the update (:759) and after it (:853-862): (R Lambda).T() & DT = Lambda.T() & DT
    const V3 LtDT = mulT(quatToMat(qOld), DT);
    const V3 Thn = old.Th + LtDT;                       // :759
    const V3 Wnew = Wold + DW;                          // :769
    st3(p.Th, c3, s, Thn);
    st3(p.W, c3, s, Wnew);
    const Q4 qNew = quatNormalize(quatMul(quatFromRotVec(DT, halfAngle(DT)), qOld));
    stq(p.Lam, c4, s, qNew); // :839-841
    x2 += dot(Wnew, Wnew) + dot(Thn, Thn);
    if (p.storeInc) { st3(p.DW, c3, s, DW); st3(p.DTh, c3, s, DT); }
    if (SCHEME == BFK_NEWMARK) {
        V3 Ac = old.Ac, dOm = old.dOm;
        if (p.first) { // predictor :167-186, and the step constants U*, Omega* (see BeamParams)
            const double k1 = p.nk1, k2 = p.nk2;
            const V3 Ao = Ac, dOo = dOm;
            const V3 Uo = ld3(p.U, c3, s) + p.gdtPrev * Ao, Oo = ld3(p.Om, c3, s) + p.gdtPrev * dOo;
            Ac = v3(-k1 * Uo.x - k2 * Ao.x, -k1 * Uo.y - k2 * Ao.y, -k1 * Uo.z - k2 * Ao.z);
            dOm = v3(-k1 * Oo.x - k2 * dOo.x, -k1 * Oo.y - k2 * dOo.y, -k1 * Oo.z - k2 * dOo.z);
            st3(p.U, c3, s, Uo + (dt * (1.0 - gamma)) * Ao);   // U = U* + gamma dt Accl from here on
            st3(p.Om, c3, s, Oo + (dt * (1.0 - gamma)) * dOo);
        }
        Ac = Ac + p.nAcc * DW;                                // :788 (and :785 through U*)
        dOm = dOm + p.nAcc * LtDT;                            // :853-862
        st3(p.Ac, c3, s, Ac); st3(p.dOm, c3, s, dOm);
    }
    if (SCHEME == BFK_EULER) { // :791-795, 864-869: U = ddt(W), Accl = ddt(U), Omega = axialVector(Lambda.T() & ddt(Lambda))
        const double idt = 1.0 / dt;
        V3 Wo, Uo;
        Q4 qo;
        if (p.first) { // the time index has just advanced: this iteration stores the old-time level
            Wo = Wold; Uo = ld3(p.U, c3, s); qo = qOld;
            st3(p.Wold, c3, s, Wo); st3(p.Uold, c3, s, Uo); stq(p.Lamold, c4, s, qo);
            st3(p.Omold, c3, s, ld3(p.Om, c3, s));
        } else {
            Wo = ld3(p.Wold, c3, s); Uo = ld3(p.Uold, c3, s); qo = ldq(p.Lamold, c4, s);
        }
        const V3 U = idt * (Wnew - Wo);
        st3(p.U, c3, s, U);
        st3(p.Ac, c3, s, idt * (U - Uo));
        const M3 Ln = quatToMat(qNew), Lo = quatToMat(qo);
        const M3 Pm = mulTM(Ln, idt * (Ln - Lo)); // Lambda.T() & ddt(Lambda)
        st3(p.Om, c3, s, v3(Pm.a[7], Pm.a[2], Pm.a[3])); // axialVector: (T.zy, T.xz, T.yx)
    }
}

// updateSolutionVariables for ONE cell (Evolve.C:759-802, 839-869): Theta, W, Lambda and the time-scheme fields.
// DW, DT = the solved increments of the cell; returns the pre-update W (the faces need it) and adds |W|^2 + |Theta|^2.
template <int SCHEME>
DEV void updateCell(const BeamParams& p, const size_t T, const int lane, const int i, const V3 DW, const V3 DT, V3& WoldOut,
                    double& x2)
{
    const CellOld old = loadCellOld<SCHEME>(p, T, lane, i);
    WoldOut = old.W;
    updateCellFrom<SCHEME>(p, T, lane, i, DW, DT, old, x2);
}

// Update of the internal face j (between cell j-1, its owner (A), and cell j, its neighbour (B)): Evolve.C:804-836, 881-913
// with the coefficients recomputed from the pre-solve state.  WoldA / WoldB: pre-update W of the two cells.
DEV void updateInternalFace(const BeamParams& p, const int b, const int j, const V3 CQ, const V3 CM, const V3 dR0, const V3 e0,
                            const double dcI, const V3 WoldA, const V3 WoldB, const V3 DWa, const V3 DTa, const V3 DWb, const V3 DTb)
{
    const size_t T = (size_t)(b >> 5);
    const int lane = b & 31, Rf = p.nmax + 1;
    const size_t f3 = tileRow(T, Rf, j, 3, lane), f4 = tileRow(T, Rf, j, 4, lane);
    const Q4 qfOld = ldq(p.Lf, f4, TS);
    const V3 Kold = ld3(p.K, f3, TS);
    const M3 LmOld = quatToMat(qfOld);
    const V3 t = dR0 + dcI * (WoldB - WoldA);
    const FaceCoef f = faceCoefficients(LmOld, CQ, CM, gammaOf(LmOld, e0, t), Kold, ld3(p.Q, f3, TS), ld3(p.M, f3, TS), t, dcI, false);
    const FaceState o = updateFace(f, qfOld, LmOld, Kold, DWa, DWb, DTa, DTb, dcI, false);
    stq(p.Lf, f4, TS, o.Lf);
    st3(p.K, f3, TS, o.K); st3(p.Q, f3, TS, o.Q); st3(p.M, f3, TS, o.M);
}
// The evaluate() of the patch fields of one end and the update of its boundary face (face 0 or face n); Wold, DW, DT
// belong to the cell next to the patch.
DEV void updatePatchFace(const BeamParams& p, const int b, const int end, const int n, const V3 CQ, const V3 CM, const double dcB,
                         const double delta, const V3 Wold, const V3 DW, const V3 DT)
{
    const size_t T = (size_t)(b >> 5);
    const int lane = b & 31, Rf = p.nmax + 1, j = end ? n : 0;
    const size_t f3 = tileRow(T, Rf, j, 3, lane), f4 = tileRow(T, Rf, j, 4, lane);
    const Q4 qfOld = ldq(p.Lf, f4, TS);
    const V3 Kold = ld3(p.K, f3, TS);
    const FaceState o = patchUpdate(p, end, b, qfOld, CQ, CM, Kold, ld3(p.Q, f3, TS), ld3(p.M, f3, TS), Wold, DW, DT, dcB, delta);
    stq(p.Lf, f4, TS, o.Lf);
    st3(p.K, f3, TS, o.K); st3(p.Q, f3, TS, o.Q); st3(p.M, f3, TS, o.M);
}
// The faces a cell owns in the update: face i+1 (internal, or the right patch for the last cell) and, for cell 0, the
// left patch (Evolve.C:804-836, 881-913 and the evaluate() of the patch fields, Theta before W).  (A) = this cell,
// (N) = cell i+1: WoldN its pre-update W, xWn / xTn its increments.
DEV void updateFacesOfCell(const BeamParams& p, const int b, const int i, const int n, const V3 CQ, const V3 CM, const V3 dR0,
                           const V3 e0, const double dcI, const double dcB, const double delta, const V3 Wold, const V3 WoldN,
                           const V3 DW, const V3 DT, const V3 xWn, const V3 xTn)
{
    if (i == n - 1) updatePatchFace(p, b, 1, n, CQ, CM, dcB, delta, Wold, DW, DT);
    else updateInternalFace(p, b, i + 1, CQ, CM, dR0, e0, dcI, Wold, WoldN, DW, DT, xWn, xTn);
    if (i == 0) updatePatchFace(p, b, 0, n, CQ, CM, dcB, delta, Wold, DW, DT);
}

template <int SCHEME>
DEV void backwardSweep(const BeamParams& p, const int b, double& dx2, double& x2)
{
    const size_t s = TS;
    const size_t T = (size_t)(b >> 5);
    const int lane = b & 31, Rc = p.nmax, Rf = p.nmax + 1;
    {
        const int n = p.nSeg[b];
        const double dZ = p.dZ[b];
        const double dcI = 1.0 / dZ, dcB = 2.0 / dZ;
        const double delta = 1.0 / dcB;
        const V3 CQ = ld3(p.CQ, b, p.Bpad), CM = ld3(p.CM, b, p.Bpad);
        V3 dR0, e0;
        {
            const M3 refL = ld9(p.refL, b, p.Bpad);
            dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
            e0 = mulT(refL, dR0);
        }
        const double dt = p.dt, beta = p.beta, gamma = p.gamma;

        V3 xWn = v3(0, 0, 0), xTn = v3(0, 0, 0); // x_{i+1}
        V3 WoldN = v3(0, 0, 0);
        for (int i = n - 1; i >= 0; i--) {
            // ---- back substitution: x_i = bPrime_i - uPrime_i & x_{i+1}
            const size_t u0 = tileRow(T, Rc, i, 36, lane), b0 = tileRow(T, Rc, i, 6, lane);
            if (i >= BEAM_PF_DIST) { // prefetch what iteration i-D reads: scratch and cell i-D, face i-D+1
                const int ic = i - BEAM_PF_DIST;
                const size_t m3 = tileRow(T, Rc, ic, 3, lane), g3 = tileRow(T, Rf, ic + 1, 3, lane);
                pfRows<36>(p.uP, tileRow(T, Rc, ic, 36, lane), s);
                pfRows<6>(p.bP, tileRow(T, Rc, ic, 6, lane), s);
                pfRows<3>(p.W, m3, s); pfRows<3>(p.Th, m3, s);
                pfRows<4>(p.Lam, tileRow(T, Rc, ic, 4, lane), s);
                if (SCHEME == BFK_NEWMARK) {
                    pfRows<3>(p.Ac, m3, s); pfRows<3>(p.dOm, m3, s);
                    if (p.first) { pfRows<3>(p.U, m3, s); pfRows<3>(p.Om, m3, s); }
                }
                pfRows<4>(p.Lf, tileRow(T, Rf, ic + 1, 4, lane), s);
                pfRows<3>(p.K, g3, s);
                pfRows<3>(p.Q, g3, s); pfRows<3>(p.M, g3, s);
            }
            double x[6];
#pragma unroll
            for (int r = 0; r < 6; r++) x[r] = ldScratch(p.bP + b0 + (size_t)r * s); // written by another SM in this launch: L2
            if (i < n - 1) {
                const double xn[6] = {xWn.x, xWn.y, xWn.z, xTn.x, xTn.y, xTn.z};
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    double acc = 0.0;
#pragma unroll
                    for (int c = 0; c < 6; c++) acc += ldScratch(p.uP + u0 + (size_t)(6 * r + c) * s) * xn[c];
                    x[r] -= acc;
                }
            }
            const V3 DW = v3(x[0], x[1], x[2]), DT = v3(x[3], x[4], x[5]);
#pragma unroll
            for (int r = 0; r < 6; r++) dx2 += x[r] * x[r];
            V3 Wold;
            updateCell<SCHEME>(p, T, lane, i, DW, DT, Wold, x2);
            updateFacesOfCell(p, b, i, n, CQ, CM, dR0, e0, dcI, dcB, delta, Wold, WoldN, DW, DT, xWn, xTn);
            xWn = DW; xTn = DT; WoldN = Wold;
        }
    }
}

// ---------------------------------------------------------------------------
// Kernels (one thread per beam).
//   beam_forward / beam_backward : the product path for bundles of at least one wave of beams: the two sweeps as two
//       launches per Newton iteration.
//   beam_newton                  : both sweeps in one launch (BEAMGPU_FUSED_KERNEL=1); measures 5-6 % slower.
// Bundles below one wave run the group-cooperative kernels of beam_coop.cuh (several lanes per beam).
// ---------------------------------------------------------------------------
DEV uint32_t warpBarrierSetup(unsigned char* smem)
{
    if ((threadIdx.x & 31) == 0) mbarInit((uint64_t*)smem + (threadIdx.x >> 5), 1);
    __syncwarp();
    return 0;
}
#define WARP_SLAB(smem, w) ((double*)((smem) + SMEM_HEADER_BYTES) + (size_t)(w) * WARP_SMEM_DOUBLES)

template <int SCHEME>
__global__ void __launch_bounds__(BEAM_THREADS, BEAM_MIN_BLOCKS) beam_newton(const __grid_constant__ BeamParams p)
{
    extern __shared__ __align__(128) unsigned char beam_smem[];
    const int b = blockIdx.x * blockDim.x + threadIdx.x, w = threadIdx.x >> 5;
    uint32_t phase = warpBarrierSetup(beam_smem);
    double v[3] = {0.0, 0.0, 0.0};
    // whole warps run the forward sweep (the prefetch is a warp-collective); padding lanes idle inside
    v[0] = forwardSweep<SCHEME>(p, b, WARP_SLAB(beam_smem, w), (uint64_t*)beam_smem + w, phase);
    if (b < p.B) backwardSweep<SCHEME>(p, b, v[1], v[2]);
    blockReduceStore<3>(v, p.partial, 0, p.nTiles, blockIdx.x);
}
template <int SCHEME>
__global__ void __launch_bounds__(BEAM_THREADS, BEAM_MIN_BLOCKS) beam_forward(const __grid_constant__ BeamParams p)
{
    extern __shared__ __align__(128) unsigned char beam_smem[];
    if (convLeave(p.cv, false)) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, w = threadIdx.x >> 5;
    uint32_t phase = warpBarrierSetup(beam_smem);
    double v[1] = {0.0};
    v[0] = forwardSweep<SCHEME>(p, b, WARP_SLAB(beam_smem, w), (uint64_t*)beam_smem + w, phase);
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.x);
}
template <int SCHEME>
__global__ void __launch_bounds__(BEAM_THREADS, BEAM_BWD_MIN_BLOCKS) beam_backward(const __grid_constant__ BeamParams p)
{
    if (convLeave(p.cv, true)) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (b < p.B) backwardSweep<SCHEME>(p, b, v[0], v[1]);
    blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.x);
}

// ---------------------------------------------------------------------------
// SPLIT sweeps ("burn at both ends", a twisted block factorisation): TWO threads per beam.
//   A bundle below one wave of beams cannot fill the GPU with one thread per beam, and an iteration then costs the latency
//   of one thread walking all n cells.  The block-Thomas elimination can start from both ends at once: the left half
//   eliminates cells 0 .. h-1 left to right exactly as forwardSweep does, the mirrored half eliminates cells n-1 .. h right to
//   left (the same code on the neighbour-side blocks, see fwdFacePart<.., MIR>), and the two meet at the face between
//   cells h-1 and h with
//       x_{h-1} = bPrime_{h-1} - uPrime_{h-1} x_h ,   x_h = bPrime_h - uPrime_h x_{h-1}
//   a 6x6 system that both halves solve (identically) at the start of the backward sweep before substituting outwards.
//   No redundant arithmetic, half the serial length, twice the threads.  Uniform bundles only (every beam nmax cells: the
//   mirrored half walks the rows downwards from nmax-1, and the bulk copies fetch one row for all 32 lanes); padding lanes
//   run as dummy beams so that a warp never leaves the straight-line path.
//   Ownership in the backward sweep: the left half updates cells 0..h-1, faces 1..h-1 and the left patch; the mirrored half
//   updates cells h..n-1, faces h..n-1 (face c together with cell c) and the right patch.  The interface face h needs the
//   pre-update W of cell h-1, which the left half overwrites concurrently: the mirrored half's forward sweep saves it (ifW).
// ---------------------------------------------------------------------------
template <int SCHEME, bool MIR, bool SAVEW>
DEV double forwardSweepSplit(const BeamParams& p, const int b, double* __restrict__ warpSm, uint64_t* bar, uint32_t& phase)
{
    const size_t s = p.Bpad;
    const int lane = threadIdx.x & 31;
    const size_t T = (size_t)(b >> 5);
    double* sm = warpSm + lane;
    double res = 0.0;
    const int n = p.nmax, hL = p.hL;
    const int m = MIR ? n - hL : hL; // cells of this part; the host selects the split sweeps only for n >= 2
#define SROW(i) (MIR ? n - 1 - (i) : (i))            // cell row of local cell i
#define SFACE(i) (MIR ? n - 1 - (i) : (i) + 1)       // the face between local cells i and i+1
    __syncwarp();
    if (m > 0) fwdPrefetchFaceRows<SCHEME>(p, T, warpSm, bar, SROW(0), SROW(1), SFACE(0));
    {
        const M3 refL = ld9(p.refL, b, s);
        const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
        smSt3(sm, R_DR0, dR0);
        smSt(sm, R_DZ, p.dZ[b]);
        smSt(sm, R_ARHO, SCHEME != BFK_STEADY ? p.ARho[b] : 0.0);
        smSt(sm, R_N, (double)m);
        smSt3(sm, R_CQ, ld3(p.CQ, b, s));
        smSt3(sm, R_CM, ld3(p.CM, b, s));
        smSt3(sm, R_CI, SCHEME != BFK_STEADY ? ld3(p.CI, b, s) : v3(0, 0, 0));
#pragma unroll
        for (int r = 0; r < 6; r++) smSt(sm, R_LB + r, 0.0);
    }
    V3 Wi = ld3(p.W, tileRow(T, p.nmax, SROW(0), 3, lane), TS);
    {
        double D36[36], s6[6];
#pragma unroll
        for (int k = 0; k < 36; k++) D36[k] = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) s6[k] = 0.0;
        patchClosure(p, MIR ? 1 : 0, b, Wi, D36, s6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            smSt(sm, R_SC + r, s6[r]);
#pragma unroll
            for (int c = 0; c < 6; c++) smSt(sm, R_DC + 6 * r + c, D36[6 * r + c]);
        }
    }
    for (int i = 0; i < m; i++) { // every cell of a half has an internal face towards the next one (the last: the interface)
        mbarWait(bar, phase & 1u);
        const FaceIn in = loadFaceIn(sm);
        if (SCHEME == BFK_STEADY) fenceProxyAsync();
        __syncwarp();
        if (SCHEME != BFK_STEADY) fwdPrefetchCell<SCHEME>(p, T, sm, SROW(i));
        else if (i + 1 < m) fwdPrefetchFaceRows<SCHEME>(p, T, warpSm, bar, SROW(i + 1), SROW(i + 2), SFACE(i + 1));
        if (MIR && i == m - 1) // pre-update W of cell h-1, for the interface face in the backward sweep
            st3(p.ifW, b, s, in.Wn);
        if (SAVEW) st3(p.wSave, tileRow(T, p.nmax, SROW(i), 3, lane), TS, Wi); // this cell's pre-update W, for beam_update
        double src[6], Dc[6][6];
        const double dz = smLd(sm, R_DZ);
        fwdFacePart<false, MIR>(p, b, sm, rcpFast(dz), 0.5 * dz, in, Dc, src, Wi);
        CellIn cell;
        if (SCHEME != BFK_STEADY) {
            cpAsyncWaitAll();
            cell = loadCellInSm<SCHEME>(sm, p.first != 0);
            fenceProxyAsync();
            __syncwarp();
            if (i + 1 < m) fwdPrefetchFaceRows<SCHEME>(p, T, warpSm, bar, SROW(i + 1), SROW(i + 2), SFACE(i + 1));
        } else
            cell = loadCellIn<SCHEME>(p, b, SROW(i));
        phase++;
        addLoadsAndInertia<SCHEME>(p, b, SROW(i), dz, smLd(sm, R_ARHO), smLd3(sm, R_CI), cell, Dc, src);
#pragma unroll
        for (int r = 0; r < 6; r++) res += src[r] * src[r];
        fwdEliminate<false>(p, b, sm, SROW(i), Dc, src);
    }
#undef SROW
#undef SFACE
    return b < p.B ? res : 0.0;
}

// x_{h-1} and x_h of a split beam from the last rows of its two halves (see above).  Pivoted 6x6 solve.
__device__ __noinline__ void splitInterface(const BeamParams& p, const size_t T, const int lane, const int h, double* __restrict__ xL,
                                            double* __restrict__ xR)
{
    const int Rc = p.nmax;
    const double* uL = p.uP + tileRow(T, Rc, h - 1, 36, lane);
    const double* uR = p.uP + tileRow(T, Rc, h, 36, lane);
    const double* bL = p.bP + tileRow(T, Rc, h - 1, 6, lane);
    const double* bR = p.bP + tileRow(T, Rc, h, 6, lane);
    double UL[6][6], UR[6][6], A[6][6], rhs[6][1], br[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        br[r] = __ldcg(bR + (size_t)r * TS);
#pragma unroll
        for (int c = 0; c < 6; c++) { UL[r][c] = __ldcg(uL + (size_t)(6 * r + c) * TS); UR[r][c] = __ldcg(uR + (size_t)(6 * r + c) * TS); }
    }
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double acc = __ldcg(bL + (size_t)r * TS);
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fma(-UL[r][c], br[c], acc);
        rhs[r][0] = acc; // bPrime_{h-1} - uPrime_{h-1} bPrime_h
#pragma unroll
        for (int c = 0; c < 6; c++) {
            double a = r == c ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) a = fma(-UL[r][k], UR[k][c], a);
            A[r][c] = a; // I - uPrime_{h-1} uPrime_h
        }
    }
    solve6<1>(A, rhs);
#pragma unroll
    for (int r = 0; r < 6; r++) xL[r] = rhs[r][0];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        double acc = br[r];
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fma(-UR[r][c], xL[c], acc);
        xR[r] = acc;
    }
}

template <int SCHEME, bool MIR>
DEV void backwardSweepSplit(const BeamParams& p, const int b, double& dx2, double& x2)
{
    const size_t s = TS;
    const size_t T = (size_t)(b >> 5);
    const int lane = b & 31, Rc = p.nmax, Rf = p.nmax + 1;
    const int n = p.nmax, hL = p.hL;
    const double dZ = p.dZ[b];
    const double dcI = 1.0 / dZ, dcB = 2.0 / dZ;
    const double delta = 1.0 / dcB;
    const V3 CQ = ld3(p.CQ, b, p.Bpad), CM = ld3(p.CM, b, p.Bpad);
    V3 dR0, e0;
    {
        const M3 refL = ld9(p.refL, b, p.Bpad);
        dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
        e0 = mulT(refL, dR0);
    }
    double xL[6], xR[6];
    splitInterface(p, T, lane, hL, xL, xR);
    // xo = x of the cell visited before this one (towards the interface), Wo = its pre-update W
    V3 xWo = MIR ? v3(xL[0], xL[1], xL[2]) : v3(xR[0], xR[1], xR[2]);
    V3 xTo = MIR ? v3(xL[3], xL[4], xL[5]) : v3(xR[3], xR[4], xR[5]);
    V3 Wo = MIR ? ld3(p.ifW, b, p.Bpad) : v3(0, 0, 0);
    const int m = MIR ? n - hL : hL;
    for (int k = 0; k < m; k++) {
        const int i = MIR ? hL + k : hL - 1 - k; // the mirrored half walks outwards to the right, the left half to the left
        {   // prefetch what the next step reads
            const int ic = MIR ? i + BEAM_PF_DIST : i - BEAM_PF_DIST;
            if (ic >= 0 && ic < n) {
                const int jf = MIR ? ic : ic + 1; // the face that cell updates
                const size_t m3 = tileRow(T, Rc, ic, 3, lane), g3 = tileRow(T, Rf, jf, 3, lane);
                pfRows<36>(p.uP, tileRow(T, Rc, ic, 36, lane), s);
                pfRows<6>(p.bP, tileRow(T, Rc, ic, 6, lane), s);
                pfRows<3>(p.W, m3, s); pfRows<3>(p.Th, m3, s);
                pfRows<4>(p.Lam, tileRow(T, Rc, ic, 4, lane), s);
                if (SCHEME == BFK_NEWMARK) {
                    pfRows<3>(p.Ac, m3, s); pfRows<3>(p.dOm, m3, s);
                    if (p.first) { pfRows<3>(p.U, m3, s); pfRows<3>(p.Om, m3, s); }
                }
                pfRows<4>(p.Lf, tileRow(T, Rf, jf, 4, lane), s);
                pfRows<3>(p.K, g3, s);
                pfRows<3>(p.Q, g3, s); pfRows<3>(p.M, g3, s);
            }
        }
        double x[6];
        if (k == 0) {
#pragma unroll
            for (int r = 0; r < 6; r++) x[r] = MIR ? xR[r] : xL[r];
        } else { // x_i = bPrime_i - uPrime_i & x_(previous cell of the walk)
            const size_t u0 = tileRow(T, Rc, i, 36, lane), b0 = tileRow(T, Rc, i, 6, lane);
            const double xn[6] = {xWo.x, xWo.y, xWo.z, xTo.x, xTo.y, xTo.z};
#pragma unroll
            for (int r = 0; r < 6; r++) {
                double acc = 0.0;
#pragma unroll
                for (int c = 0; c < 6; c++) acc += ldScratch(p.uP + u0 + (size_t)(6 * r + c) * s) * xn[c];
                x[r] = ldScratch(p.bP + b0 + (size_t)r * s) - acc;
            }
        }
        const V3 DW = v3(x[0], x[1], x[2]), DT = v3(x[3], x[4], x[5]);
#pragma unroll
        for (int r = 0; r < 6; r++) dx2 += x[r] * x[r];
        V3 Wold;
        updateCell<SCHEME>(p, T, lane, i, DW, DT, Wold, x2);
        if (MIR) { // face i: owner = cell i-1 (the previous cell of the walk), neighbour = this cell
            updateInternalFace(p, b, i, CQ, CM, dR0, e0, dcI, Wo, Wold, xWo, xTo, DW, DT);
            if (i == n - 1) updatePatchFace(p, b, 1, n, CQ, CM, dcB, delta, Wold, DW, DT);
        } else {   // face i+1: owner = this cell, neighbour = the previous cell of the walk; the interface face is the other half's
            if (k > 0) updateInternalFace(p, b, i + 1, CQ, CM, dR0, e0, dcI, Wold, Wo, DW, DT, xWo, xTo);
            if (i == 0) updatePatchFace(p, b, 0, n, CQ, CM, dcB, delta, Wold, DW, DT);
        }
        xWo = DW; xTo = DT; Wo = Wold;
    }
}

// grid (ceil(B / BEAM_THREADS), 2): blockIdx.y = 0 the left halves, 1 the mirrored halves
template <int SCHEME, bool SAVEW>
__global__ void __launch_bounds__(BEAM_THREADS, BEAM_MIN_BLOCKS) beam_forward_split(const __grid_constant__ BeamParams p)
{
    extern __shared__ __align__(128) unsigned char beam_smem[];
    if (convLeave(p.cv, false)) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, w = threadIdx.x >> 5;
    uint32_t phase = warpBarrierSetup(beam_smem);
    double v[1] = {0.0};
    if ((b & ~31) < p.Bpad) { // a warp wholly past the padded bundle has no tile in memory (lanes B..Bpad-1 run as dummy beams)
        if (blockIdx.y) v[0] = forwardSweepSplit<SCHEME, true, SAVEW>(p, b, WARP_SLAB(beam_smem, w), (uint64_t*)beam_smem + w, phase);
        else v[0] = forwardSweepSplit<SCHEME, false, SAVEW>(p, b, WARP_SLAB(beam_smem, w), (uint64_t*)beam_smem + w, phase);
    }
    blockReduceStore<1>(v, p.partial, 0, p.nTiles, blockIdx.y * gridDim.x + blockIdx.x);
}
template <int SCHEME>
__global__ void __launch_bounds__(BEAM_THREADS, BEAM_BWD_MIN_BLOCKS) beam_backward_split(const __grid_constant__ BeamParams p)
{
    if (convLeave(p.cv, true)) return;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    double v[2] = {0.0, 0.0};
    if (b < p.B) {
        if (blockIdx.y) backwardSweepSplit<SCHEME, true>(p, b, v[0], v[1]);
        else backwardSweepSplit<SCHEME, false>(p, b, v[0], v[1]);
    }
    blockReduceStore<2>(v, p.partial, 1, p.nTiles, blockIdx.y * gridDim.x + blockIdx.x);
}

// Source contribution of the point forces for every cell of every beam (one thread per beam, its cells in turn).
__global__ void beam_point_forces(const __grid_constant__ BeamParams p)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int n = p.nSeg[b];
    for (int i = 0; i < n; i++) {
        double o[6];
        pointForceSource(p, b, i, o);
        const size_t c6 = tileRow((size_t)(b >> 5), p.nmax, i, 6, b & 31);
#pragma unroll
        for (int r = 0; r < 6; r++) p.pfSrc[c6 + (size_t)r * TS] = o[r];
    }
}

// beamEnergyData (functionObjects/beamEnergyData/beamEnergyData.C:84-215): per beam the strain energy of the face
// strains Gamma, K (0.5 L[own] on internal faces, 0.25 L[cell] on the end faces) and the linear / angular kinetic
// energy of the cells, straight from the device state.  One thread per beam; out[3 b + k].
template <int SCHEME>
__global__ void beam_energy(const __grid_constant__ BeamParams p, double* __restrict__ out)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const size_t T = (size_t)(b >> 5), sb = p.Bpad;
    const int lane = b & 31, n = p.nSeg[b], Rc = p.nmax, Rf = p.nmax + 1;
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ;
    const V3 CQ = ld3(p.CQ, b, sb), CM = ld3(p.CM, b, sb);
    const M3 refL = ld9(p.refL, b, sb);
    const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]), e0 = mulT(refL, dR0);
    auto strain = [&](int j, V3 t, V3 e) { // Gamma.CQ.Gamma + K.CM.K of face j
        const V3 G = gammaOf(quatToMat(ldq(p.Lf, tileRow(T, Rf, j, 4, lane), TS)), e, t);
        const V3 K = ld3(p.K, tileRow(T, Rf, j, 3, lane), TS);
        return dot(G, cmul(CQ, G)) + dot(K, cmul(CM, K));
    };
    auto cellLen = [&](int i) { return p.Lc ? p.Lc[tileRow(T, Rc, i, 1, lane)] : dZ; };
    double intE = 0.0, kinL = 0.0, kinA = 0.0;
    V3 Wprev = ld3(p.W, tileRow(T, Rc, 0, 3, lane), TS);
    intE += 0.25 * cellLen(0) * strain(0, dcB * (ld3(p.Wb, b, sb) - Wprev) - dR0, -e0);
    for (int i = 0; i < n; i++) {
        const double Li = cellLen(i);
        if (i + 1 < n) {
            const V3 Wn = ld3(p.W, tileRow(T, Rc, i + 1, 3, lane), TS);
            intE += 0.5 * Li * strain(i + 1, dR0 + dcI * (Wn - Wprev), e0);
            Wprev = Wn;
        } else
            intE += 0.25 * Li * strain(n, dR0 + dcB * (ld3(p.Wb, (size_t)3 * sb + b, sb) - Wprev), e0);
        if (SCHEME != BFK_STEADY) {
            const size_t c3 = tileRow(T, Rc, i, 3, lane);
            V3 U = ld3(p.U, c3, TS), Om = ld3(p.Om, c3, TS);
            if (SCHEME == BFK_NEWMARK) { // stored as step constants (BeamParams): U = U* + gamma dt Accl
                U = U + p.gdtPrev * ld3(p.Ac, c3, TS);
                Om = Om + p.gdtPrev * ld3(p.dOm, c3, TS);
            }
            kinL += 0.5 * Li * p.ARho[b] * dot(U, U);
            kinA += 0.5 * Li * dot(Om, cmul(ld3(p.CI, b, sb), Om));
        }
    }
    out[3 * (size_t)b] = intE; out[3 * (size_t)b + 1] = kinL; out[3 * (size_t)b + 2] = kinA;
}

// Deterministic final reduction of the per-block partial sums (fixed order).  partial[k * stride + block]: slot 0 is
// written by the nF blocks of the forward launch, slots 1 and 2 by the nB blocks of the backward launch (the two
// launches of an iteration may come from different kernel families, see bf_ctx::coopFwd / coopBwd).
// One block of RED_THREADS threads; every thread adds its strided share of each slot with four independent accumulators (the
// cell-parallel backward part leaves one partial per 128 cells -- 12,800 of them for 8192 beams: 256 threads adding 50 values
// one dependent L2 round trip after the other made this kernel 40 us of a 630 us iteration).
#define RED_THREADS 1024
DEV void reducePartialsBlock(const double* __restrict__ partial, int stride, int nF, int nB, double (&sm)[3][RED_THREADS])
{
    const int t = threadIdx.x, nt = blockDim.x;
    {
        const double* src = partial;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        int i = t;
        for (; i + 3 * nt < nF; i += 4 * nt) { a0 += src[i]; a1 += src[i + nt]; a2 += src[i + 2 * nt]; a3 += src[i + 3 * nt]; }
        for (; i < nF; i += nt) a0 += src[i];
        sm[0][t] = (a0 + a1) + (a2 + a3);
    }
    { // the two slots of the backward part side by side (eight loads in flight per thread; each slot in the order above)
        const double *s1 = partial + (size_t)stride, *s2 = partial + 2 * (size_t)stride;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0, b3 = 0.0;
        int i = t;
        for (; i + 3 * nt < nB; i += 4 * nt) {
            a0 += s1[i]; a1 += s1[i + nt]; a2 += s1[i + 2 * nt]; a3 += s1[i + 3 * nt];
            b0 += s2[i]; b1 += s2[i + nt]; b2 += s2[i + 2 * nt]; b3 += s2[i + 3 * nt];
        }
        for (; i < nB; i += nt) { a0 += s1[i]; b0 += s2[i]; }
        sm[1][t] = (a0 + a1) + (a2 + a3);
        sm[2][t] = (b0 + b1) + (b2 + b3);
    }
    __syncthreads();
    for (int o = nt / 2; o > 0; o >>= 1) {
        if (t < o)
            for (int k = 0; k < 3; k++) sm[k][t] += sm[k][t + o];
        __syncthreads();
    }
}
__global__ void __launch_bounds__(RED_THREADS) beam_reduce_partials(const double* __restrict__ partial, int stride, int nF, int nB,
                                                                    double* __restrict__ out, double* __restrict__ outHost)
{
    __shared__ double sm[3][RED_THREADS];
    reducePartialsBlock(partial, stride, nF, nB, sm);
    if (threadIdx.x < 3) {
        out[threadIdx.x] = sm[threadIdx.x][0];
        if (outHost) outHost[threadIdx.x] = sm[threadIdx.x][0]; // mapped pinned host memory
    }
}
struct PeerTable {
    double* bank[CONV_MAX_RANKS]; // the current bank of every rank's exchange buffer (own included)
};
// The same reduction, then this rank's record of iteration cv.iter into every rank's buffer.
__global__ void __launch_bounds__(RED_THREADS) beam_reduce_exchange(const double* __restrict__ partial, int stride, int nF, int nB,
                                                                    const ConvCtl cv, const PeerTable peers, const int rank)
{
    __shared__ double sm[3][RED_THREADS];
    __shared__ int ended;
    if (threadIdx.x >= RED_THREADS - 32) { // the last warp (the first ones carry the tail of the tree): did the sweeps of this iteration run?
        const int d = cv.iter > 0 ? convDecide(cv, cv.iter - 1, true) : 0;
        if (threadIdx.x == RED_THREADS - 32) ended = d == 1;
    }
    reducePartialsBlock(partial, stride, nF, nB, sm);
    if ((int)threadIdx.x < cv.nRanks) {
        double* rec = peers.bank[threadIdx.x] + ((size_t)cv.iter * cv.nRanks + rank) * 4;
        rec[0] = sm[0][0]; rec[1] = sm[1][0]; rec[2] = sm[2][0];
        __threadfence_system();
        stReleaseSys(rec + 3, cv.tag + (ended ? 1.0 : 0.0));
    }
}
// Global sums and verdict of iterations k0 .. k1-1 into mapped host memory, [k - k0][4]: sums, then 1 = the loop ends with
// this iteration, 2 = it had ended before, 0 = it goes on.  Launched after the last enqueued iteration of a chunk.
__global__ void beam_publish_chunk(const ConvCtl cv, const int k0, const int k1, double* __restrict__ outHost)
{
    const int k = k0 + (int)(threadIdx.x >> 5); // one warp per iteration of the chunk
    if (k >= k1) return;
    double s[4];
    const int ended = convSums(cv, k, true, s);
    if ((threadIdx.x & 31) == 0) {
        double* o = outHost + 4 * (k - k0);
        o[0] = s[0]; o[1] = s[1]; o[2] = s[2];
        o[3] = ended == 2 ? 3.0 : ended ? 2.0 : convRules(cv, k, s) ? 1.0 : 0.0; // 3: a peer's record never arrived
        __threadfence_system();
    }
}
__global__ void beam_publish_sums(const double* __restrict__ in, double* __restrict__ outHost)
{
    if (threadIdx.x < 3) outHost[threadIdx.x] = in[threadIdx.x];
}
