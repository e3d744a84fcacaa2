// beamgpu.cu — C-ABI (include/beamgpu.h) over the fused CUDA kernels.  sm_100a only.
//
// Host side of libbeamgpu.so: owns the device-resident state in the beam-interleaved SoA
// layout, transposes to/from the reference's array-of-structs ordering at the boundary
// (bf_set_field / bf_get_field), and drives the Newton loop of evolve()
// (coupledTotalLagNewtonRaphsonBeamEvolve.C:55-670) with checkConvergence (:926-1026) on the
// host, exactly where the reference evaluates it.  There is no CPU compute path here.
#include "../../include/beamgpu.h"
#include "beam_kernels.cuh"
#include "beam_long.cuh"
#include "beam_coop.cuh"
#include "beam_bicgstab.cuh"
#include "beam_contrib.cuh"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define BF_GREAT 1e15
#define THREADS BEAM_THREADS
#define PIPE_CHUNK_MAX 16

static char g_create_err[512] = "";

struct Id128 { // ncclUniqueId: 128 opaque bytes, passed by value to ncclCommInitRank
    char internal[128];
};
struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
struct PipeState { // the pipelined Newton loop of one evolve() (evolvePipelined; bf_evolve_begin / bf_evolve_end)
    bool active;   // between bf_evolve_begin and bf_evolve_end
    bool deferred; // bf_evolve_begin could not enqueue ahead (host-driven loop): bf_evolve_end runs bf_evolve
    bool done;
    bf_tolerances tol;
    ConvCtl cv;
    int k0, k1, chunk, evSet, nIter, status;
    double initialResidualNorm, currentResidualNorm, deltaXNorm, XNorm;
};
struct bf_ctx {
    char err[512];
    PipeState pipe;
    int device;
    cudaStream_t stream;
    int64_t B, N, F;
    int Bpad, nmax, grid;
    // bundles below one wave of beams: group-cooperative kernels (beam_coop.cuh), selectable per sweep
    int coopFwd, coopBwd, coopGrid;
    int wsBpb; // beams (half beams with split) per block of beam_forward_ws: 16 (two eliminator lanes per beam) or 32 (one)
    int wsEw;  // eliminator warps per block: 1, or 2 (wsBpb = 32: the two column groups of a beam in two warps)
    int hL;    // split sweeps: cells of the left part of every beam (BeamParams::hL)
    int split; // two threads per beam (beam_forward_split / beam_backward_split): uniform bundles between the two other families
    double* d_ifW;
    double* d_X; // coopBwd: the solution increments, tiled [tile][nmax][6][32] (beam_backsub -> beam_update)
    double* d_Wsave; // coopBwd: the pre-update W of every cell, tiled like W (saved by the forward part, read by beam_update)
    int partialStride; // leading dimension of d_partial: the largest block count of any kernel family
    std::vector<int32_t> nSeg, hostWType; // hostWType: [2][B]
    std::vector<int64_t> cellStart, faceStart;
    bool uniformN;
    // device pools
    void* pools[32];
    int nPools;
    int32_t *d_nSeg, *d_wType, *d_thType;
    int64_t *d_cellStart, *d_faceStart;
    double *d_dZ, *d_CQ, *d_CM, *d_ARho, *d_CI, *d_weight, *d_refL;
    double *d_W, *d_Th, *d_Lam, *d_U, *d_Ac, *d_Om, *d_dOm, *d_DW, *d_DTh, *d_q, *d_m, *d_Lc;
    double *d_Wold, *d_Uold, *d_Omold, *d_Lamold; // Euler d2dt2Scheme: old-time level
    double *d_pf, *d_pfSrc;                       // point forces, 9 per cell, and their source terms, 6 per cell (on first use)
    bool hasPf, hasWeight;
    ContribParams contrib; // beamMomentumContribution list (bf_set_momentum_contributions); contrib.n == 0: none
    double *d_Lf, *d_K, *d_Q, *d_M; // d_Lf, d_Lam: unit quaternions (4 components); Gamma is derived, not stored
    double *d_Wb, *d_Thb, *d_Lamb, *d_DWb, *d_DThb, *d_force, *d_wPresc, *d_thPresc, *d_follower, *d_offset,
        *d_moment, *d_springK, *d_springDir;
    double *d_uP, *d_bP, *d_partial, *d_sums;
    // few long beams: one thread per cell, assembled block rows, block parallel cyclic reduction (beam_long.cuh)
    int longMode, longBlocks;
    int32_t* d_cellBeam;
    double* d_longRows[12]; // L, D, U, R: the assembled rows and two working copies
    double* d_longX;
    double* d_stage; // staging buffer in host (AoS) ordering
    double* d_stagePatch; // separate small staging for patch values (never shared with a pending mirror)
    double* d_stageOut;   // staging of the step outputs (bf_set_step_outputs): [2][B][3]
    int outEnd;
    double *outW, *outTheta;
    double* d_stageQ;     // host-ordered quaternions (4 per cell/face) between the transposition and the tensor conversion
    double* d_stagePatchQ;
    size_t stageDoubles;
    double* h_sums;  // pinned + mapped: the kernels write the norms straight into host memory, so the
                     // per-iteration read never queues behind a bulk D2H copy on the copy engine
    double* d_hsums; // device alias of h_sums
    // options / time state
    int scheme, storeInc;
    double betaN, gammaN, deltaT;
    double dtStar; // the time step the stored Newmark constants U*, Omega* belong to (beam_kernels.cuh, BeamParams)
    int iOuterCorr;
    bool timeAdvanced; // set by bf_begin_step, cleared by the first Newton iteration after it: the Newmark predictor and the
                       // Euler old-time capture belong to the time index, not to a call of evolve() (Evolve.C:167-186 reads oldTime())
    bool hasQ, hasM;
    bf_reduce_fn reducer;
    void* reducerUser;
    // NCCL
    NcclApi nccl;
    void* comm;
    // coupled bundles (beam_bicgstab.cuh): adjacency of the couplings, factors of the tridiagonal part, Krylov vectors
    int64_t nPairs;
    int64_t *d_adjPtr, *d_adjCol;
    double *d_adjK, *d_cplDinv, *d_cplUP, *d_cplVec; // d_cplVec: r, rw, p, ph, v, s, sh, t (8 x 6N)
    double *h_lin, *d_hlin;                         // pinned + mapped: dot products of the Krylov loop
    int linMaxIter, linIterations;
    double linTol, linRelTol;
    // pipelined Newton loop (device-side convergence test, one-sided exchange of the norm sums; beam_kernels.cuh ConvCtl)
    double* d_xbuf;            // [2 banks][CONV_MAXK][xRanks][4], cudaMalloc (its IPC handle is exported to the peers)
    double* d_vflag;           // [2 banks][CONV_MAXK] verdict cache (ConvCtl::vflag), local
    unsigned long long peerWaitNs; // ConvCtl::waitNs: 20 s, or BEAMGPU_PEER_TIMEOUT_S seconds
    double* peerXbuf[CONV_MAX_RANKS]; // every rank's buffer as mapped into this process (own: d_xbuf)
    int xRanks, xRank;
    long long epoch;
    int lastIters;             // Newton iterations of the previous step: how far ahead the next one is enqueued
    double* h_res;             // pinned + mapped: verdicts of a chunk of iterations, [PIPE_CHUNK_MAX][4]
    double* d_hres;
    ConvCtl cvNow;             // filled by bf_evolve for launchIteration; xbuf == nullptr: the host decides
    // per enqueued iteration of a chunk: start, between the sweeps, end (timing).  Two sets: the elapsed times of a chunk are
    // read AFTER the next chunk (normally the next time step) has been enqueued, while the GPU works -- a dozen
    // cudaEventElapsedTime calls between two steps are 1 % of a 1.6 ms step of a small bundle
    cudaEvent_t evIter[2][16][3];
    int evSet;                 // the set the next chunk records into
    int evPendingSet, evPendingK0, evPendingN; // a finished chunk whose times have not been read yet (evPendingN iterations ran)
    // the 1-block reduction of iteration k runs on a side stream while the (speculative) forward part of iteration k+1 already
    // occupies the GPU: only the backward part of k+1 needs the verdict.  The partial sums alternate between two banks.
    cudaStream_t redStream;
    cudaEvent_t evSwept, evReduced;
    bool redPending;           // a reduction on redStream that the main stream has not waited for yet
    int updateWarpBlocks;      // beam_update with one warp per block (BEAMGPU_UPDATE_WARP_BLOCKS = 0: four)
    int sideReduce;            // BEAMGPU_NO_SIDE_REDUCE = 1: the reduction stays in line on the main stream
    // instrumentation
    int64_t launches;
    int timing, splitKernels;
    cudaEvent_t evRegion[2];
    // asynchronous mirroring (bf_mirror_begin / bf_mirror_wait)
    cudaStream_t copyStream;
    cudaEvent_t evStaged, evCopied;
    bool copyPending;
    cudaEvent_t ev[3];
    double fwdMs, bwdMs;
    int64_t timedIters;
    double fwdMsK[8], bwdMsK[8], gapMsK[8]; // BEAMGPU_TIMING_DUMP: the same by iteration index of the step (gap: end of iteration k-1 to start of k)
    int64_t itersK[8];
};

static int fail(bf_ctx* c, int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(c ? c->err : g_create_err, 512, fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_OK(c, call)                                                                        \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(c, BF_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                        __FILE__, __LINE__);                                                    \
    } while (0)

extern "C" const char* bf_version(void) { return "beamgpu 0.1 (sm_100a)"; }
extern "C" const char* bf_backend(void) { return "cuda-sm100a"; }
extern "C" const char* bf_last_error(const bf_ctx* c) { return c ? c->err : g_create_err; }
extern "C" int64_t bf_num_cells(const bf_ctx* c) { return c->N; }
extern "C" int64_t bf_num_internal_faces(const bf_ctx* c) { return c->F; }
extern "C" int64_t bf_num_beams(const bf_ctx* c) { return c->B; }
extern "C" int64_t bf_launch_count(const bf_ctx* c) { return c->launches; }
extern "C" int32_t bf_split_point(const bf_ctx* c) { return c->split ? c->hL : 0; }
extern "C" int32_t bf_kernel_family(const bf_ctx* c) { return (c->longMode || c->nPairs) ? BF_FAMILY_CELL : (c->split && c->coopFwd) ? BF_FAMILY_SPLIT_GROUP : (c->split && c->coopBwd) ? BF_FAMILY_SPLIT_CELL : c->split ? BF_FAMILY_SPLIT : c->coopFwd ? BF_FAMILY_GROUP : BF_FAMILY_PER_BEAM; }

extern "C" int bf_host_alloc(void** p, int64_t bytes)
{
    return cudaHostAlloc(p, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? BF_OK : BF_ERR_NOMEM;
}
extern "C" int bf_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? BF_OK : BF_ERR_CUDA; }

// ---------------------------------------------------------------------------
// layout transposition: reference ordering (AoS, beam-major) <-> warp-tiled [b/32][row][c][b%32]
// ---------------------------------------------------------------------------
__device__ __forceinline__ int findBeam(const int64_t* __restrict__ start, int B, int64_t g)
{
    int lo = 0, hi = B; // start[lo] <= g < start[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (start[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// cells: host index g = cellStart[b] + i ;  dev index (((b/32)*R + i)*nc + c)*32 + b%32, R rows per beam
// faces: host index g = faceStart[b] + (j-1), j = 1..n-1 ; dev face index j = i + faceOffset
// Tiled through shared memory so that BOTH sides are coalesced: a block moves a tile of
// 32 beams x 32 cells x nc components; device accesses run along beams (threadIdx.x = beam),
// host-order accesses run along the contiguous (cell, component) run of one beam.
template <bool TO_DEVICE>
__global__ void transposeInternal(double* __restrict__ dev, double* __restrict__ host, const int64_t* __restrict__ start,
                                  int B, int R, int nc, int faceOffset)
{
    extern __shared__ double tile[]; // [32 cells][nc][33]
    const int b0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y; // 32 x 8
    const int run = 32 * nc;                      // doubles of one beam inside the tile (host order)
    if (!TO_DEVICE) {
        const int b = b0 + tx;
        const int nb = b < B ? (int)(start[b + 1] - start[b]) : 0;
        for (int k = ty; k < run; k += 8) { // k = il*nc + c
            const int il = k / nc, c = k - il * nc;
            const int i = i0 + il;
            if (i < nb) tile[(size_t)k * 33 + tx] = dev[(((size_t)blockIdx.x * R + (i + faceOffset)) * nc + c) * 32 + tx];
        }
        __syncthreads();
        for (int bl = ty; bl < 32; bl += 8) {
            const int bb = b0 + bl;
            if (bb >= B) continue;
            const int64_t s0 = start[bb];
            const int nbb = (int)(start[bb + 1] - s0);
            for (int k = tx; k < run; k += 32) {
                const int il = k / nc;
                if (i0 + il < nbb) host[(s0 + i0) * nc + k] = tile[(size_t)k * 33 + bl];
            }
        }
    } else {
        for (int bl = ty; bl < 32; bl += 8) {
            const int bb = b0 + bl;
            if (bb >= B) continue;
            const int64_t s0 = start[bb];
            const int nbb = (int)(start[bb + 1] - s0);
            for (int k = tx; k < run; k += 32) {
                const int il = k / nc;
                if (i0 + il < nbb) tile[(size_t)k * 33 + bl] = host[(s0 + i0) * nc + k];
            }
        }
        __syncthreads();
        const int b = b0 + tx;
        const int nb = b < B ? (int)(start[b + 1] - start[b]) : 0;
        for (int k = ty; k < run; k += 8) {
            const int il = k / nc, c = k - il * nc;
            const int i = i0 + il;
            if (i < nb) dev[(((size_t)blockIdx.x * R + (i + faceOffset)) * nc + c) * 32 + tx] = tile[(size_t)k * 33 + tx];
        }
    }
}
// patch values: host [b*nc + c] ; dev end arrays ((end*nc + c)*Bpad + b), or for face arrays the
// warp-tiled element of the patch face (face 0 or nSeg[b]; R = nmax+1 rows per beam)
template <bool TO_DEVICE>
__global__ void transposePatch(double* __restrict__ dev, double* __restrict__ host, const int32_t* __restrict__ nSeg,
                               int B, int Bpad, int R, int nc, int end, int faceArray)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    for (int c = 0; c < nc; c++) {
        const size_t d = faceArray ? ((((size_t)(b >> 5) * R + (end ? nSeg[b] : 0)) * nc + c) * 32 + (b & 31))
                                   : (((size_t)end * nc + c) * Bpad + b);
        if (TO_DEVICE) dev[d] = host[(size_t)b * nc + c]; else host[(size_t)b * nc + c] = dev[d];
    }
}
// rows x nc x width array of identity rotations: tensors (nc = 9: xx = yy = zz = 1) or quaternions (nc = 4: w = 1)
__global__ void fillIdentity(double* __restrict__ T, int64_t rows, int width, int nc)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rows * width) return;
    const int64_t r = k / width, b = k % width;
    T[(r * nc + 0) * width + b] = 1.0;
    if (nc == 9) {
        T[(r * 9 + 4) * width + b] = 1.0;
        T[(r * 9 + 8) * width + b] = 1.0;
    }
}

// Lambdaf = I at start (CTL.C:289-314): the stored product Lambdaf & refLambdaf starts as refLambdaf.
// One thread per (tile row, lane) of the warp-tiled face array [tile][R][4][32].
__global__ void fillRefRotation(double* __restrict__ Lf, int64_t rowsTotal, int R, const double* __restrict__ refL, int Bpad)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= rowsTotal * 32) return;
    const int64_t row = k / 32;
    const int lane = (int)(k % 32);
    const int b = (int)(row / R) * 32 + lane;
    const Q4 q = matToQuat(ld9(refL, b, Bpad));
    double* o = Lf + row * 4 * 32 + lane;
    o[0] = q.w; o[32] = q.x; o[64] = q.y; o[96] = q.z;
}

// host-ordered rotation tensors (9) <-> unit quaternions (4): the device keeps Lambda as a quaternion and, for the
// face field, the quaternion of Lambdaf & refLambdaf (refL != nullptr; start = faceStart to find the beam of an
// internal face, nullptr for patch values, where element g belongs to beam g).
template <bool TO_QUAT>
__global__ void convertRotation(double* __restrict__ tens, double* __restrict__ quat, int64_t count,
                                const double* __restrict__ refL, const int64_t* __restrict__ start, int B, int Bpad)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= count) return;
    M3 ref = m3eye();
    if (refL) {
        const int b = start ? findBeam(start, B, g) : (int)g;
        ref = ld9(refL, b, Bpad);
    }
    if (TO_QUAT) {
        M3 R;
        for (int k = 0; k < 9; k++) R.a[k] = tens[g * 9 + k];
        const Q4 q = matToQuat(refL ? mul(R, ref) : R);
        quat[g * 4] = q.w; quat[g * 4 + 1] = q.x; quat[g * 4 + 2] = q.y; quat[g * 4 + 3] = q.z;
    } else {
        M3 R = quatToMat(q4(quat[g * 4], quat[g * 4 + 1], quat[g * 4 + 2], quat[g * 4 + 3]));
        if (refL) R = mulBT(R, ref);
        for (int k = 0; k < 9; k++) tens[g * 9 + k] = R.a[k];
    }
}
// Gamma = refLambdaf.T() & ((Lambdaf.T() & (dR0Ds + snGrad(W))) - dR0Ds)  (Evolve.C:881-887), evaluated from the
// stored W / Lambdaf when the host asks for the field.  One thread per beam; internal faces into host order,
// patch faces into left[b*3..], right[b*3..].
__global__ void gammaField(const __grid_constant__ BeamParams p, const int64_t* __restrict__ faceStart, double* __restrict__ internal,
                           double* __restrict__ left, double* __restrict__ right)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const size_t T = (size_t)(b >> 5);
    const int lane = b & 31, n = p.nSeg[b], Rc = p.nmax, Rf = p.nmax + 1;
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ;
    const M3 refL = ld9(p.refL, b, p.Bpad);
    const V3 dR0 = v3(refL.a[0], refL.a[3], refL.a[6]);
    const V3 e0 = mulT(refL, dR0);
    V3 Wprev = ld3(p.W, tileRow(T, Rc, 0, 3, lane), TS);
    if (left) {
        const V3 g = gammaOf(quatToMat(ldq(p.Lf, tileRow(T, Rf, 0, 4, lane), TS)), -e0,
                             dcB * (ld3(p.Wb, b, p.Bpad) - Wprev) - dR0);
        left[3 * (size_t)b] = g.x; left[3 * (size_t)b + 1] = g.y; left[3 * (size_t)b + 2] = g.z;
    }
    for (int j = 1; j < n; j++) {
        const V3 Wn = ld3(p.W, tileRow(T, Rc, j, 3, lane), TS);
        if (internal) {
            const V3 g = gammaOf(quatToMat(ldq(p.Lf, tileRow(T, Rf, j, 4, lane), TS)), e0, dR0 + dcI * (Wn - Wprev));
            double* o = internal + 3 * (faceStart[b] + j - 1);
            o[0] = g.x; o[1] = g.y; o[2] = g.z;
        }
        Wprev = Wn;
    }
    if (right) {
        const V3 g = gammaOf(quatToMat(ldq(p.Lf, tileRow(T, Rf, n, 4, lane), TS)), e0,
                             dR0 + dcB * (ld3(p.Wb, (size_t)3 * p.Bpad + b, p.Bpad) - Wprev));
        right[3 * (size_t)b] = g.x; right[3 * (size_t)b + 1] = g.y; right[3 * (size_t)b + 2] = g.z;
    }
}

struct bf_ctx;
template <bool TO_DEVICE>
static void launchTranspose(bf_ctx* c, double* dev, double* hostOrder, int nc, int surface);

template <typename T>
static T* carve(char*& cur, size_t count)
{
    T* p = (T*)cur;
    cur += ((count * sizeof(T) + 255) / 256) * 256;
    return p;
}

template <bool TO_DEVICE>
static void launchTranspose(bf_ctx* c, double* dev, double* hostOrder, int nc, int surface)
{
    const int rows = surface ? c->nmax - 1 : c->nmax; // internal faces per beam = cells - 1
    if (rows <= 0) return;
    const dim3 grid((unsigned)((c->B + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
    const size_t smem = sizeof(double) * 32 * nc * 33;
    transposeInternal<TO_DEVICE><<<grid, block, smem, c->stream>>>(
        dev, hostOrder, surface ? c->d_faceStart : c->d_cellStart, (int)c->B, surface ? c->nmax + 1 : c->nmax, nc, surface);
}

static void readChunkTimes(bf_ctx* c);
extern "C" int bf_destroy(bf_ctx* c)
{
    if (!c) return BF_OK;
    cudaSetDevice(c->device);
    if (c->comm && c->nccl.CommDestroy) c->nccl.CommDestroy(c->comm);
    readChunkTimes(c);
    if (getenv("BEAMGPU_TIMING_DUMP"))
        for (int k = 0; k < 8; k++)
            if (c->itersK[k]) fprintf(stderr, "beamgpu rank %d iteration %d of a step: fwd %.4f ms  bwd %.4f ms  gap before %.4f ms  (%lld samples)\n", c->xRank, k,
                                      c->fwdMsK[k] / c->itersK[k], c->bwdMsK[k] / c->itersK[k], c->gapMsK[k] / c->itersK[k], (long long)c->itersK[k]);
    for (int i = 0; i < c->nPools; i++) cudaFree(c->pools[i]);
    if (c->h_sums) cudaFreeHost(c->h_sums);
    if (c->h_res) cudaFreeHost(c->h_res);
    if (c->h_lin) cudaFreeHost(c->h_lin);
    for (int r = 0; r < c->xRanks; r++)
        if (r != c->xRank && c->peerXbuf[r]) cudaIpcCloseMemHandle(c->peerXbuf[r]);
    if (c->d_xbuf) cudaFree(c->d_xbuf);
    if (c->d_vflag) cudaFree(c->d_vflag);
    for (int s = 0; s < 2; s++) for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) if (c->evIter[s][i][k]) cudaEventDestroy(c->evIter[s][i][k]);
    for (int i = 0; i < 3; i++) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    for (int i = 0; i < 2; i++) if (c->evRegion[i]) cudaEventDestroy(c->evRegion[i]);
    if (c->redStream) { cudaStreamSynchronize(c->redStream); cudaStreamDestroy(c->redStream); }
    if (c->evSwept) cudaEventDestroy(c->evSwept);
    if (c->evReduced) cudaEventDestroy(c->evReduced);
    if (c->copyStream) { cudaStreamSynchronize(c->copyStream); cudaStreamDestroy(c->copyStream); }
    if (c->evStaged) cudaEventDestroy(c->evStaged);
    if (c->evCopied) cudaEventDestroy(c->evCopied);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return BF_OK;
}

static int devAlloc(bf_ctx* c, void** p, size_t bytes)
{
    if (c->nPools >= (int)(sizeof c->pools / sizeof c->pools[0])) return fail(c, BF_ERR_NOMEM, "device pool table full");
    CUDA_OK(c, cudaMalloc(p, bytes));
    CUDA_OK(c, cudaMemsetAsync(*p, 0, bytes, c->stream));
    c->pools[c->nPools++] = *p;
    return BF_OK;
}

// Buffers of the cell-parallel path (beam_long.cuh): the assembled block rows, two working copies, the solution and the
// beam of every cell.  Allocated by bf_create for few long beams, on first use by bf_newton_iteration_checked otherwise.
static int ensureLongBuffers(bf_ctx* c)
{
    if (c->d_cellBeam) return BF_OK;
    char* cur;
    const size_t N = (size_t)c->N;
    if (int rc = devAlloc(c, (void**)&cur, sizeof(double) * N * (3 * (36 * 3 + 6) + 6) + sizeof(int32_t) * N + 256 * 16)) return rc;
    for (int k = 0; k < 12; k++) c->d_longRows[k] = carve<double>(cur, (k % 4 == 3 ? 6 : 36) * N);
    c->d_longX = carve<double>(cur, 6 * N + 6); // one row of slack: the last cell reads "x of the next cell" unguarded
    c->d_cellBeam = carve<int32_t>(cur, N);
    std::vector<int32_t> cb(N);
    for (int64_t b = 0; b < c->B; b++) for (int64_t g = c->cellStart[b]; g < c->cellStart[b + 1]; g++) cb[g] = (int32_t)b;
    CUDA_OK(c, cudaMemcpyAsync(c->d_cellBeam, cb.data(), sizeof(int32_t) * N, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}

// Replaces the constructor state set-up of CTL.C:74-1182 for what the hot path touches.
extern "C" int bf_create(const bf_layout* lay, const bf_beam_props* pr, const bf_bc* bc, const bf_options* opt,
                         bf_ctx** out)
{
    if (!lay || !pr || !bc || !opt || !out || lay->nBeams <= 0 || !lay->nSeg || !lay->length || !pr->CQ || !pr->CM ||
        !bc->wType || !bc->thetaType)
        return fail(nullptr, BF_ERR_INVALID, "bf_create: null/invalid argument");
    if (opt->scheme != BF_SCHEME_STEADY && opt->scheme != BF_SCHEME_NEWMARK && opt->scheme != BF_SCHEME_EULER)
        return fail(nullptr, BF_ERR_UNSUPPORTED, "bf_create: unknown d2dt2 scheme %d (steadyState, Newmark, Euler)",
                    opt->scheme);
    if (lay->nBeams > (int64_t)1 << 30) return fail(nullptr, BF_ERR_INVALID, "bf_create: too many beams");
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
        return fail(nullptr, BF_ERR_CUDA, "bf_create: no CUDA device visible (libbeamgpu has no CPU fallback)");
    if (opt->device < 0 || opt->device >= nDev)
        return fail(nullptr, BF_ERR_INVALID, "bf_create: device %d out of range (%d visible)", opt->device, nDev);

    bf_ctx* c = new bf_ctx();
    memset(c->err, 0, sizeof c->err);
    c->device = opt->device;
    c->stream = nullptr; c->nPools = 0; c->h_sums = nullptr; c->comm = nullptr;
    c->ev[0] = c->ev[1] = c->ev[2] = nullptr;
    c->reducer = nullptr; c->reducerUser = nullptr;
    c->launches = 0; c->timing = 0; c->splitKernels = 1;
    c->redStream = nullptr; c->evSwept = c->evReduced = nullptr; c->redPending = false;
    memset(&c->pipe, 0, sizeof c->pipe);
    c->updateWarpBlocks = 1;
    c->sideReduce = getenv("BEAMGPU_NO_SIDE_REDUCE") ? 0 : 1;
    if (const char* e = getenv("BEAMGPU_UPDATE_WARP_BLOCKS")) c->updateWarpBlocks = atoi(e) ? 1 : 0;
    // Product path: the two sweeps as two launches per Newton iteration (beam_forward, beam_backward).  The single fused
    // launch (beam_newton: every thread runs both sweeps) measures 5-6 % slower on B200: with all SMs in the same phase
    // the latency-bound forward sweep is not disturbed by the HBM-saturating backward sweep of a neighbouring block.
    // BEAMGPU_FUSED_KERNEL=1 (or BEAMGPU_SPLIT_KERNELS=0) selects the fused launch.
    if (const char* e = getenv("BEAMGPU_FUSED_KERNEL")) c->splitKernels = atoi(e) ? 0 : 1;
    if (const char* e = getenv("BEAMGPU_SPLIT_KERNELS")) c->splitKernels = atoi(e) ? 1 : 0;
    c->evRegion[0] = c->evRegion[1] = nullptr;
    c->nPairs = 0; c->d_adjPtr = c->d_adjCol = nullptr; c->d_adjK = c->d_cplDinv = c->d_cplUP = c->d_cplVec = nullptr;
    c->h_lin = c->d_hlin = nullptr; c->linMaxIter = 1000; c->linIterations = 0; c->linTol = 1e-13; c->linRelTol = 0.0;
    c->peerWaitNs = 20000000000ull;
    if (const char* e = getenv("BEAMGPU_PEER_TIMEOUT_S")) { const double t = atof(e); if (t > 0) c->peerWaitNs = (unsigned long long)(t * 1e9); }
    c->d_stageOut = nullptr; c->outEnd = 0; c->outW = c->outTheta = nullptr;
    c->d_xbuf = nullptr; c->d_vflag = nullptr; c->xRanks = 1; c->xRank = 0; c->epoch = 0; c->lastIters = 4; c->h_res = nullptr; c->d_hres = nullptr;
    memset(&c->cvNow, 0, sizeof c->cvNow);
    for (int i = 0; i < CONV_MAX_RANKS; i++) c->peerXbuf[i] = nullptr;
    for (int s = 0; s < 2; s++) for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) c->evIter[s][i][k] = nullptr;
    c->evSet = 0; c->evPendingN = 0; c->evPendingSet = 0; c->evPendingK0 = 0;
    c->copyStream = nullptr; c->evStaged = c->evCopied = nullptr; c->copyPending = false; c->fwdMs = c->bwdMs = 0; c->timedIters = 0;
    memset(c->fwdMsK, 0, sizeof c->fwdMsK); memset(c->bwdMsK, 0, sizeof c->bwdMsK); memset(c->gapMsK, 0, sizeof c->gapMsK); memset(c->itersK, 0, sizeof c->itersK);
    c->hasQ = c->hasM = false;
    c->d_q = c->d_m = c->d_Lc = c->d_DW = c->d_DTh = nullptr;
    c->d_Wold = c->d_Uold = c->d_Omold = c->d_Lamold = c->d_pf = c->d_pfSrc = nullptr;
    c->hasPf = false;
#define CREATE_FAIL(code, ...) do { int rc_ = fail(nullptr, code, __VA_ARGS__); bf_destroy(c); return rc_; } while (0)
#define CREATE_CUDA(call)                                                                          \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) CREATE_FAIL(BF_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    } while (0)
    CREATE_CUDA(cudaSetDevice(c->device));
    // the sweeps park their face tensors in dynamic shared memory (> 48 KB needs the opt-in)
    CREATE_CUDA(cudaFuncSetAttribute(transposeInternal<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 9 * 33 * 8));
    CREATE_CUDA(cudaFuncSetAttribute(transposeInternal<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 9 * 33 * 8));
#define OPT_IN_SMEM(KERNEL)                                                                                               \
    CREATE_CUDA(cudaFuncSetAttribute(KERNEL<BFK_STEADY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));  \
    CREATE_CUDA(cudaFuncSetAttribute(KERNEL<BFK_NEWMARK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)); \
    CREATE_CUDA(cudaFuncSetAttribute(KERNEL<BFK_EULER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES))
    OPT_IN_SMEM(beam_newton);
    OPT_IN_SMEM(beam_forward);
#define OPT_IN_SMEM2(KERNEL, FLAG)                                                                                                \
    CREATE_CUDA((cudaFuncSetAttribute(KERNEL<BFK_STEADY, FLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)));  \
    CREATE_CUDA((cudaFuncSetAttribute(KERNEL<BFK_NEWMARK, FLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES))); \
    CREATE_CUDA((cudaFuncSetAttribute(KERNEL<BFK_EULER, FLAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)))
    OPT_IN_SMEM2(beam_forward_split, false); OPT_IN_SMEM2(beam_forward_split, true);
#define OPT_IN_WS(BPB, SPLIT, EW)                                                                                                                              \
    CREATE_CUDA((cudaFuncSetAttribute(beam_forward_ws<BFK_STEADY, BPB, SPLIT, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES(BPB, EW))));  \
    CREATE_CUDA((cudaFuncSetAttribute(beam_forward_ws<BFK_NEWMARK, BPB, SPLIT, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES(BPB, EW)))); \
    CREATE_CUDA((cudaFuncSetAttribute(beam_forward_ws<BFK_EULER, BPB, SPLIT, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES(BPB, EW))))
    OPT_IN_WS(16, false, 1); OPT_IN_WS(16, true, 1); OPT_IN_WS(32, false, 1); OPT_IN_WS(32, true, 1); OPT_IN_WS(32, false, 2); OPT_IN_WS(32, true, 2);
    CREATE_CUDA(cudaFuncSetAttribute(beam_backsub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BS_SMEM_BYTES));
    CREATE_CUDA(cudaFuncSetAttribute(beam_backsub_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)BSS_SMEM_BYTES));
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->redStream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaEventCreateWithFlags(&c->evSwept, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&c->evReduced, cudaEventDisableTiming));
    for (int i = 0; i < 3; i++) CREATE_CUDA(cudaEventCreate(&c->ev[i]));
    for (int i = 0; i < 2; i++) CREATE_CUDA(cudaEventCreate(&c->evRegion[i]));
    CREATE_CUDA(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
    CREATE_CUDA(cudaEventCreateWithFlags(&c->evStaged, cudaEventDisableTiming));
    CREATE_CUDA(cudaEventCreateWithFlags(&c->evCopied, cudaEventDisableTiming));

    const int64_t B = c->B = lay->nBeams;
    c->nSeg.assign(lay->nSeg, lay->nSeg + B);
    c->cellStart.assign(B + 1, 0);
    c->faceStart.assign(B + 1, 0);
    int nmax = 0;
    bool uniform = true;
    for (int64_t b = 0; b < B; b++) {
        if (lay->nSeg[b] < 1 || !(lay->length[b] > 0)) CREATE_FAIL(BF_ERR_INVALID, "bf_create: beam %lld has nSeg<1 or length<=0", (long long)b);
        c->cellStart[b + 1] = c->cellStart[b] + lay->nSeg[b];
        c->faceStart[b + 1] = c->faceStart[b] + lay->nSeg[b] - 1;
        if (lay->nSeg[b] > nmax) nmax = lay->nSeg[b];
        if (lay->nSeg[b] != lay->nSeg[0]) uniform = false;
    }
    c->uniformN = uniform;
    c->N = c->cellStart[B];
    c->F = c->faceStart[B];
    c->nmax = nmax;
    const int Bpad = c->Bpad = (int)((B + 31) / 32 * 32);
    c->grid = (int)((B + THREADS - 1) / THREADS);
    c->scheme = opt->scheme;
    c->storeInc = opt->storeIncrements ? 1 : 0;
    c->betaN = opt->newmarkBeta > 0 ? opt->newmarkBeta : 0.25;
    c->gammaN = opt->newmarkGamma > 0 ? opt->newmarkGamma : 0.5;
    c->deltaT = 1.0;
    c->dtStar = 1.0;
    c->iOuterCorr = 0;
    c->timeAdvanced = true;
    const bool newmark = c->scheme != BF_SCHEME_STEADY; // U, Accl, Omega, dotOmega exist for both transient schemes
    const bool euler = c->scheme == BF_SCHEME_EULER;
    const size_t bp = (size_t)Bpad, nc_ = (size_t)nmax, nf_ = (size_t)nmax + 1;

    // ---- per-beam + end arrays (one pool)
    {
        const size_t bytes = 256 * 44 + sizeof(double) * bp * (3 + 1 + 3 + 3 + 1 + 3 + 3 + 9 + 2 * (3 + 3 + 9 + 3 + 3 + 3 + 3 + 3 + 3 + 3 + 3 + 1 + 3)) +
                             sizeof(int32_t) * bp * 5 + sizeof(int64_t) * 2 * (size_t)(B + 1);
        char* cur;
        if (devAlloc(c, (void**)&cur, bytes)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
        c->d_dZ = carve<double>(cur, bp); c->d_CQ = carve<double>(cur, 3 * bp); c->d_CM = carve<double>(cur, 3 * bp);
        c->d_ARho = carve<double>(cur, bp); c->d_CI = carve<double>(cur, 3 * bp); c->d_weight = carve<double>(cur, 3 * bp);
        c->d_refL = carve<double>(cur, 9 * bp);
        c->d_Wb = carve<double>(cur, 6 * bp); c->d_Thb = carve<double>(cur, 6 * bp); c->d_Lamb = carve<double>(cur, 18 * bp);
        c->d_DWb = carve<double>(cur, 6 * bp); c->d_DThb = carve<double>(cur, 6 * bp); c->d_force = carve<double>(cur, 6 * bp);
        c->d_wPresc = carve<double>(cur, 6 * bp); c->d_thPresc = carve<double>(cur, 6 * bp);
        c->d_follower = carve<double>(cur, 6 * bp); c->d_offset = carve<double>(cur, 6 * bp);
        c->d_moment = carve<double>(cur, 6 * bp); c->d_springK = carve<double>(cur, 2 * bp); c->d_ifW = carve<double>(cur, 3 * bp);
        c->d_springDir = carve<double>(cur, 6 * bp);
        c->d_nSeg = carve<int32_t>(cur, bp); c->d_wType = carve<int32_t>(cur, 2 * bp); c->d_thType = carve<int32_t>(cur, 2 * bp);
        c->d_cellStart = carve<int64_t>(cur, (size_t)B + 1); c->d_faceStart = carve<int64_t>(cur, (size_t)B + 1);
    }
    // ---- cell / face state
    {
        const size_t cellD = nc_ * bp * (3 + 3 + 4 + (newmark ? 12 : 0) + (euler ? 13 : 0) + (c->storeInc ? 6 : 0));
        const size_t faceD = nf_ * bp * (4 + 3 + 3 + 3);
        char* cur;
        if (devAlloc(c, (void**)&cur, sizeof(double) * (cellD + faceD) + 256 * 24)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
        c->d_W = carve<double>(cur, 3 * nc_ * bp); c->d_Th = carve<double>(cur, 3 * nc_ * bp); c->d_Lam = carve<double>(cur, 4 * nc_ * bp);
        if (newmark) {
            c->d_U = carve<double>(cur, 3 * nc_ * bp); c->d_Ac = carve<double>(cur, 3 * nc_ * bp);
            c->d_Om = carve<double>(cur, 3 * nc_ * bp); c->d_dOm = carve<double>(cur, 3 * nc_ * bp);
        } else {
            c->d_U = c->d_Ac = c->d_Om = c->d_dOm = nullptr;
        }
        if (euler) {
            c->d_Wold = carve<double>(cur, 3 * nc_ * bp); c->d_Uold = carve<double>(cur, 3 * nc_ * bp);
            c->d_Omold = carve<double>(cur, 3 * nc_ * bp); c->d_Lamold = carve<double>(cur, 4 * nc_ * bp);
        }
        if (c->storeInc) { c->d_DW = carve<double>(cur, 3 * nc_ * bp); c->d_DTh = carve<double>(cur, 3 * nc_ * bp); }
        c->d_Lf = carve<double>(cur, 4 * nf_ * bp); c->d_K = carve<double>(cur, 3 * nf_ * bp);
        c->d_Q = carve<double>(cur, 3 * nf_ * bp); c->d_M = carve<double>(cur, 3 * nf_ * bp);
    }
    // ---- block-Thomas scratch + partial sums
    {
        char* cur;
        c->wsBpb = 16; c->wsEw = 1;
        c->coopGrid = (Bpad + 15) / 16;
        c->partialStride = 2 * c->grid; // the split sweeps launch two blocks per 128 beams
        if ((int)((c->N + 127) / 128) > c->partialStride) c->partialStride = (int)((c->N + 127) / 128);
        if (2 * c->coopGrid > c->partialStride) c->partialStride = 2 * c->coopGrid;
        const int cellGrid = (int)(((size_t)Bpad * (size_t)nmax + 31) / 32); // beam_update: one thread per (padded) cell, one warp per block
        if (cellGrid > c->partialStride) c->partialStride = cellGrid;
        if (devAlloc(c, (void**)&cur, sizeof(double) * (42 * nc_ * bp + 6 * (size_t)c->partialStride + 8) + 256 * 6)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
        c->d_uP = carve<double>(cur, 36 * nc_ * bp); c->d_bP = carve<double>(cur, 6 * nc_ * bp);
        c->d_partial = carve<double>(cur, 6 * (size_t)c->partialStride); c->d_sums = carve<double>(cur, 8); // partial sums: two banks (pipelined loop: by iteration parity)
        // Kernel family by bundle size (measured on B200, 200-cell Newmark beams, ms per Newton iteration):
        //      beams        4096   8192   16384  32768  65536
        //      per beam     1.17   1.18   1.20   1.84   3.69   one thread per beam: flat below one wave (37,888 beams)
        //      split        0.61   0.63   1.01   1.93   3.51   two threads per beam, half the serial length
        //      split cell   0.45   0.61   1.17                 split forward sweep, backward part with one thread per cell
        //      split group  0.395  0.76   1.41                assembler + eliminator warps per half beam (16 per block), cell-parallel backward
        //      group        0.60   0.96   1.87   3.48   6.72   the same for whole beams (serves ragged bundles)
        // With w = B/32 warps and S = 8 warps x SMs resident, one thread per beam costs ceil(w/S) sweep rounds and the split
        // sweeps ceil(2w/S)/2: split when that is fewer, or equal with more than one round (the shorter tasks fill the last
        // round better).  The split sweeps need a uniform bundle; small ragged bundles take the group kernels.  The split group
        // kernels occupy 8 warps per 32 beams: they serve bundles whose B/4 warps are resident at once.
        // BEAMGPU_SPLIT = 1 / 0, BEAMGPU_COOP = 1 / 0 / fwd / bwd and BEAMGPU_WS = 16 / 32 / 33 force a family (tests, scripts/gpu_sizes.py).
        {
            int nSm = 148;
            cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, c->device);
            const long S = 8L * nSm, w = (B + 31) / 32;
            const long r1 = 2 * ((w + S - 1) / S), r2 = (2 * w + S - 1) / S; // in half rounds
            c->split = (r2 < r1 || (r2 == r1 && w > S)) ? 1 : 0;
            if (const char* e = getenv("BEAMGPU_SPLIT")) c->split = atoi(e) ? 1 : 0;
            if (!c->uniformN || nmax < 2) c->split = 0;
            c->coopFwd = c->coopBwd = (!c->split && B < 6144) ? 1 : 0;
            if (const char* e = getenv("BEAMGPU_COOP")) {
                if (!strcmp(e, "fwd")) { c->coopFwd = 1; c->coopBwd = 0; }
                else if (!strcmp(e, "bwd")) { c->coopFwd = 0; c->coopBwd = 1; }
                else c->coopFwd = c->coopBwd = atoi(e) ? 1 : 0;
                if (c->coopFwd && !getenv("BEAMGPU_WS")) c->split = 0;
            }
            // BEAMGPU_WS = 16 / 32: the warp-specialised forward kernel with that many beams per block (with BEAMGPU_SPLIT=1: per
            // half beam); 33: 32 per block with two eliminator warps
            if (const char* e = getenv("BEAMGPU_WS")) {
                c->wsBpb = atoi(e) >= 32 ? 32 : 16;
                c->wsEw = atoi(e) == 33 ? 2 : 1;
                c->coopFwd = 1;
            }
            if (c->split && !getenv("BEAMGPU_COOP") && !getenv("BEAMGPU_WS")) {
                c->coopBwd = B < 9216 ? 1 : 0;                  // cell-parallel backward part: ahead below ~5k beams, level with the split sweep up to 8192
                c->coopFwd = 8L * w <= S ? 1 : 0;               // split group forward: all its warps resident at once
            }
        }
        // Where the two threads of a beam meet.  Equal halves for the single-wave families; the split SWEEPS of a bundle that
        // takes several rounds of resident blocks (65,536 beams: 512 + 512 blocks of 100 cells on 296 slots = 3.46 rounds, the
        // last one half empty) choose the parts so that the rounds come out even: blocks are dispatched in order (all left parts,
        // then all mirrored parts), so list scheduling of `grid` tasks of hL cells followed by `grid` tasks of nmax - hL cells on
        // the resident slots predicts the length of the launch in cell steps; e.g. 150 + 50 cells -> 350 steps instead of 400.
        c->hL = (nmax + 1) / 2;
        if (c->split && !c->coopFwd && !c->coopBwd && nmax >= 8) {
            int nSm = 148;
            cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, c->device);
            const int slots = BEAM_MIN_BLOCKS * nSm, nb = c->grid;
            auto makespan = [&](int a) {
                std::vector<long> h(slots, 0); // min-heap of slot finish times
                auto cmp = [](long x, long y) { return x > y; };
                for (int pass = 0; pass < 2; pass++)
                    for (int t = 0; t < nb; t++) {
                        std::pop_heap(h.begin(), h.end(), cmp);
                        h.back() += pass ? nmax - a : a;
                        std::push_heap(h.begin(), h.end(), cmp);
                    }
                return *std::max_element(h.begin(), h.end());
            };
            const long equal = makespan(c->hL);
            long best = equal;
            for (int a = c->hL + 1; a <= nmax - 2; a++) {
                const long m_ = makespan(a);
                if (m_ < best) { best = m_; c->hL = a; }
            }
            if (best * 100 > equal * 97) c->hL = (nmax + 1) / 2; // not worth leaving the symmetric split
        }
        if (const char* e = getenv("BEAMGPU_SPLIT_AT")) { const int a = atoi(e); if (a >= 1 && a < nmax) c->hL = a; }
        c->d_X = c->d_Wsave = nullptr;
        if (c->coopBwd && devAlloc(c, (void**)&c->d_X, sizeof(double) * 9 * nc_ * bp)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
        if (c->coopBwd) c->d_Wsave = c->d_X + 6 * nc_ * bp;
        // BEAMGPU_LONG_BEAMS = 1 / 0 forces the cell-parallel path on / off; by default it serves bundles that cannot
        // occupy the GPU with one thread per beam but have enough cells per beam to do so with one thread per cell
        const char* lm = getenv("BEAMGPU_LONG_BEAMS");
        c->longMode = lm ? atoi(lm) != 0 : (nmax >= 512 && B <= 128);
        c->longBlocks = (int)((c->N + 127) / 128);
        c->d_cellBeam = nullptr; c->d_longX = nullptr;
        for (int k = 0; k < 12; k++) c->d_longRows[k] = nullptr;
    }
    // ---- staging buffer (largest field in host ordering: a tensor vol field)
    c->stageDoubles = (size_t)c->N * 9;
    if (devAlloc(c, (void**)&c->d_stage, sizeof(double) * c->stageDoubles)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
    if (devAlloc(c, (void**)&c->d_stagePatch, sizeof(double) * (size_t)B * 9)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
    if (devAlloc(c, (void**)&c->d_stageQ, sizeof(double) * (size_t)c->N * 4)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
    if (devAlloc(c, (void**)&c->d_stagePatchQ, sizeof(double) * (size_t)B * 4)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
    CREATE_CUDA(cudaHostAlloc((void**)&c->h_sums, 8 * sizeof(double), cudaHostAllocMapped));
    CREATE_CUDA(cudaHostGetDevicePointer((void**)&c->d_hsums, c->h_sums, 0));
    CREATE_CUDA(cudaHostAlloc((void**)&c->h_res, PIPE_CHUNK_MAX * 4 * sizeof(double), cudaHostAllocMapped));
    CREATE_CUDA(cudaHostGetDevicePointer((void**)&c->d_hres, c->h_res, 0));
    for (int s = 0; s < 2; s++) for (int i = 0; i < 16; i++) for (int k = 0; k < 3; k++) CREATE_CUDA(cudaEventCreate(&c->evIter[s][i][k]));
    // the exchange buffer is its own cudaMalloc: CUDA IPC exports whole allocations
    CREATE_CUDA(cudaMalloc((void**)&c->d_xbuf, sizeof(double) * 2 * CONV_MAXK * CONV_MAX_RANKS * 4));
    CREATE_CUDA(cudaMemsetAsync(c->d_xbuf, 0, sizeof(double) * 2 * CONV_MAXK * CONV_MAX_RANKS * 4, c->stream));
    c->peerXbuf[0] = c->d_xbuf;
    CREATE_CUDA(cudaMalloc((void**)&c->d_vflag, sizeof(double) * 2 * CONV_MAXK));
    CREATE_CUDA(cudaMemsetAsync(c->d_vflag, 0, sizeof(double) * 2 * CONV_MAXK, c->stream));

    // ---- upload constants
    {
        std::vector<double> h(bp * 9);
        std::vector<int32_t> hi(bp * 2);
        auto up = [&](double* dst, size_t n) { return cudaMemcpyAsync(dst, h.data(), n * sizeof(double), cudaMemcpyHostToDevice, c->stream); };
        std::fill(h.begin(), h.end(), 1.0);
        for (int64_t b = 0; b < B; b++) h[b] = lay->length[b] / lay->nSeg[b]; // createBeamMesh.C:245
        CREATE_CUDA(up(c->d_dZ, bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        auto fill3 = [&](const double* src, double dflt) {
            std::fill(h.begin(), h.end(), dflt);
            if (src) for (int64_t b = 0; b < B; b++) for (int k = 0; k < 3; k++) h[k * bp + b] = src[3 * b + k];
        };
        fill3(pr->CQ, 1.0); CREATE_CUDA(up(c->d_CQ, 3 * bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        fill3(pr->CM, 1.0); CREATE_CUDA(up(c->d_CM, 3 * bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        fill3(pr->CIRho, 0.0); CREATE_CUDA(up(c->d_CI, 3 * bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        fill3(pr->weight, 0.0);
        c->hasWeight = false;
        for (double w : h) c->hasWeight = c->hasWeight || w != 0.0;
        CREATE_CUDA(up(c->d_weight, 3 * bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        std::fill(h.begin(), h.end(), 0.0);
        if (pr->ARho) for (int64_t b = 0; b < B; b++) h[b] = pr->ARho[b];
        CREATE_CUDA(up(c->d_ARho, bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        std::fill(h.begin(), h.end(), 0.0);
        for (size_t b = 0; b < bp; b++) { h[0 * bp + b] = 1.0; h[4 * bp + b] = 1.0; h[8 * bp + b] = 1.0; }
        if (pr->refLambdaf) for (int64_t b = 0; b < B; b++) {
            // setInitialPositionBeam writes a rotation; rounding in its file I/O is removed here (Gram-Schmidt on the
            // columns) because the sweeps use refLambdaf.T() & refLambdaf = I exactly
            double R[9];
            for (int k = 0; k < 9; k++) R[k] = pr->refLambdaf[9 * b + k];
            double c0[3] = {R[0], R[3], R[6]}, c1[3] = {R[1], R[4], R[7]};
            const double n0 = sqrt(c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2]);
            if (!(n0 > 0)) CREATE_FAIL(BF_ERR_INVALID, "bf_create: refLambdaf of beam %lld is singular", (long long)b);
            for (int k = 0; k < 3; k++) c0[k] /= n0;
            const double d01 = c0[0] * c1[0] + c0[1] * c1[1] + c0[2] * c1[2];
            for (int k = 0; k < 3; k++) c1[k] -= d01 * c0[k];
            const double n1 = sqrt(c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2]);
            if (!(n1 > 0)) CREATE_FAIL(BF_ERR_INVALID, "bf_create: refLambdaf of beam %lld is singular", (long long)b);
            for (int k = 0; k < 3; k++) c1[k] /= n1;
            const double c2[3] = {c0[1] * c1[2] - c0[2] * c1[1], c0[2] * c1[0] - c0[0] * c1[2], c0[0] * c1[1] - c0[1] * c1[0]};
            const double Rn[9] = {c0[0], c1[0], c2[0], c0[1], c1[1], c2[1], c0[2], c1[2], c2[2]};
            for (int k = 0; k < 9; k++) h[k * bp + b] = Rn[k];
        }
        CREATE_CUDA(up(c->d_refL, 9 * bp)); CREATE_CUDA(cudaStreamSynchronize(c->stream));
        // boundary-condition classes and spring data
        std::vector<double> sk(2 * bp, 0.0), sd(6 * bp, 0.0);
        c->hostWType.assign(2 * (size_t)B, 0);
        for (int end = 0; end < 2; end++)
            for (int64_t b = 0; b < B; b++) {
                const int64_t pI = end * B + b;
                const int wt = bc->wType[pI], tt = bc->thetaType[pI];
                if (wt < BF_W_FIXED || wt > BF_W_SPRING || tt < BF_TH_FIXED || tt > BF_TH_MOMENT)
                    CREATE_FAIL(BF_ERR_INVALID, "bf_create: unknown boundary-condition class at patch %lld", (long long)pI);
                hi[end * bp + b] = wt;
                c->hostWType[(size_t)end * B + b] = wt;
                if (wt == BF_W_SPRING) {
                    if (!bc->springStiffness || !bc->springDirection)
                        CREATE_FAIL(BF_ERR_INVALID, "bf_create: spring BC without linearSpringCoeffs");
                    double d[3] = {bc->springDirection[3 * pI], bc->springDirection[3 * pI + 1], bc->springDirection[3 * pI + 2]};
                    const double mg = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
                    if (mg > BF_SMALL) for (int k = 0; k < 3; k++) d[k] /= mg; // linearSpringForce...C:116-119
                    sk[end * bp + b] = bc->springStiffness[pI];
                    for (int k = 0; k < 3; k++) sd[((size_t)end * 3 + k) * bp + b] = d[k];
                }
            }
        CREATE_CUDA(cudaMemcpyAsync(c->d_wType, hi.data(), 2 * bp * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaStreamSynchronize(c->stream));
        for (int end = 0; end < 2; end++) for (int64_t b = 0; b < B; b++) hi[end * bp + b] = bc->thetaType[end * B + b];
        CREATE_CUDA(cudaMemcpyAsync(c->d_thType, hi.data(), 2 * bp * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaMemcpyAsync(c->d_springK, sk.data(), 2 * bp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaMemcpyAsync(c->d_springDir, sd.data(), 6 * bp * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        std::vector<int32_t> ns(bp, 1);
        for (int64_t b = 0; b < B; b++) ns[b] = lay->nSeg[b];
        CREATE_CUDA(cudaMemcpyAsync(c->d_nSeg, ns.data(), bp * sizeof(int32_t), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaMemcpyAsync(c->d_cellStart, c->cellStart.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaMemcpyAsync(c->d_faceStart, c->faceStart.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, c->stream));
        CREATE_CUDA(cudaStreamSynchronize(c->stream));
    }
    // Lambda = Lambdaf = I (CTL.C:289-314), also the patch values of the vol field Lambda
    {
        // tiled arrays: (Bpad/32)*R rows of 32 lanes
        int64_t rows = (int64_t)nmax * (Bpad / 32);
        fillIdentity<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_Lam, rows, 32, 4);
        rows = ((int64_t)nmax + 1) * (Bpad / 32);
        fillRefRotation<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, c->stream>>>(c->d_Lf, rows, nmax + 1, c->d_refL, Bpad);
        fillIdentity<<<(unsigned)((2 * Bpad + 255) / 256), 256, 0, c->stream>>>(c->d_Lamb, 2, Bpad, 9);
        c->launches += 3;
        CREATE_CUDA(cudaGetLastError());
    }
    if (c->longMode && ensureLongBuffers(c)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
    if (lay->cellLength) { // L() per cell (HermiteSpline::segLengths), otherwise dZ
        char* cur;
        if (devAlloc(c, (void**)&cur, sizeof(double) * nc_ * bp)) { int rc_ = fail(nullptr, BF_ERR_NOMEM, "%s", c->err); bf_destroy(c); return rc_; }
        c->d_Lc = (double*)cur;
        CREATE_CUDA(cudaMemcpyAsync(c->d_stage, lay->cellLength, sizeof(double) * c->N, cudaMemcpyHostToDevice, c->stream));
        launchTranspose<true>(c, c->d_Lc, c->d_stage, 1, 0);
        c->launches++;
    }
    CREATE_CUDA(cudaStreamSynchronize(c->stream));
    *out = c;
    return BF_OK;
}

// ---------------------------------------------------------------------------
// How a field is represented on the device.  FK_STAR: a Newmark velocity stored as its step constant
// (U* = U - gamma dt Accl), FK_RATE: the rate such a constant refers to (setting it must keep the velocity).
enum { FK_PLAIN = 0, FK_ROTATION = 1, FK_GAMMA = 2, FK_STAR = 3, FK_RATE = 4 };
static double* starPartner(bf_ctx* c, double* dev)
{
    if (dev == c->d_U) return c->d_Ac;
    if (dev == c->d_Om) return c->d_dOm;
    if (dev == c->d_Ac) return c->d_U;
    if (dev == c->d_dOm) return c->d_Om;
    return nullptr;
}
// out = a + g * x over a whole tiled cell array (3 components)
__global__ void axpyCells(double* __restrict__ out, const double* __restrict__ a, const double* __restrict__ x, double g, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = a[k] + g * x[k];
}
// star += g * (rateOld - rateNew); rateOld = rateNew   (the velocity star + g * rate is unchanged)
__global__ void rebaseStar(double* __restrict__ star, double* __restrict__ rate, const double* __restrict__ rateNew, double g, size_t n)
{
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { star[k] += g * (rate[k] - rateNew[k]); rate[k] = rateNew[k]; }
}
static int fieldInfo(bf_ctx* c, int field, double** dev, double** endArr, int* nc, int* surface, int* kind)
{
    *endArr = nullptr; *surface = 0; *nc = 3; *kind = FK_PLAIN; *dev = nullptr;
    switch (field) {
    case BF_FIELD_W: *dev = c->d_W; *endArr = c->d_Wb; break;
    case BF_FIELD_THETA: *dev = c->d_Th; *endArr = c->d_Thb; break;
    case BF_FIELD_LAMBDA: *dev = c->d_Lam; *endArr = c->d_Lamb; *nc = 9; *kind = FK_ROTATION; break;
    case BF_FIELD_DW: *dev = c->d_DW; *endArr = c->d_DWb; break;
    case BF_FIELD_DTHETA: *dev = c->d_DTh; *endArr = c->d_DThb; break;
    // the step-constant representation belongs to the Newmark scheme; Euler keeps the plain fields
    case BF_FIELD_U: *dev = c->d_U; *kind = c->scheme == BF_SCHEME_NEWMARK ? FK_STAR : FK_PLAIN; break;
    case BF_FIELD_ACCL: *dev = c->d_Ac; *kind = c->scheme == BF_SCHEME_NEWMARK ? FK_RATE : FK_PLAIN; break;
    case BF_FIELD_OMEGA: *dev = c->d_Om; *kind = c->scheme == BF_SCHEME_NEWMARK ? FK_STAR : FK_PLAIN; break;
    case BF_FIELD_DOTOMEGA: *dev = c->d_dOm; *kind = c->scheme == BF_SCHEME_NEWMARK ? FK_RATE : FK_PLAIN; break;
    case BF_FIELD_LAMBDAF: *dev = c->d_Lf; *nc = 9; *surface = 1; *kind = FK_ROTATION; break;
    case BF_FIELD_GAMMA: *dev = c->d_Lf; *surface = 1; *kind = FK_GAMMA; break;
    case BF_FIELD_K: *dev = c->d_K; *surface = 1; break;
    case BF_FIELD_Q: *dev = c->d_Q; *surface = 1; break;
    case BF_FIELD_M: *dev = c->d_M; *surface = 1; break;
    default: return fail(c, BF_ERR_INVALID, "unknown field id %d", field);
    }
    return BF_OK;
}

static int drainMirror(bf_ctx* c)
{
    if (c->copyPending) { CUDA_OK(c, cudaStreamSynchronize(c->copyStream)); c->copyPending = false; }
    return BF_OK;
}
static BeamParams makeParams(bf_ctx* c);
static int ensureLongBuffers(bf_ctx* c);
static int solveCoupled(bf_ctx* c, const CoupledParams& q);
// Device state -> host-ordered (createBeamMesh) array `dst` in device memory, on the compute stream.
static void stageInternal(bf_ctx* c, double* dev, int nc, int surface, int kind, double* dst)
{
    const int64_t count = surface ? c->F : c->N;
    if (count == 0) return;
    if (kind == FK_PLAIN || kind == FK_RATE) {
        launchTranspose<false>(c, dev, dst, nc, surface);
        c->launches++;
    } else if (kind == FK_STAR) { // velocity = constant + gamma dt * rate, formed in the (idle) block-Thomas scratch
        const size_t n = (size_t)3 * c->nmax * c->Bpad;
        axpyCells<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_uP, dev, starPartner(c, dev), c->gammaN * c->dtStar, n);
        launchTranspose<false>(c, c->d_uP, dst, nc, surface);
        c->launches += 2;
    } else if (kind == FK_ROTATION) {
        launchTranspose<false>(c, dev, c->d_stageQ, 4, surface);
        convertRotation<false><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(
            dst, c->d_stageQ, count, surface ? c->d_refL : nullptr, c->d_faceStart, (int)c->B, c->Bpad);
        c->launches += 2;
    } else {
        gammaField<<<(unsigned)((c->B + 127) / 128), 128, 0, c->stream>>>(makeParams(c), c->d_faceStart, dst, nullptr, nullptr);
        c->launches++;
    }
}
static int xferInternal(bf_ctx* c, double* dev, double* host, int nc, int surface, int kind, bool toDevice)
{
    if (int rc = drainMirror(c)) return rc;
    const int64_t count = surface ? c->F : c->N;
    if (count == 0) return BF_OK;
    if (toDevice) {
        if (kind == FK_GAMMA) return BF_OK; // derived from W and Lambdaf on the device (Evolve.C:881-887): nothing to store
        CUDA_OK(c, cudaMemcpyAsync(c->d_stage, host, sizeof(double) * count * nc, cudaMemcpyHostToDevice, c->stream));
        if (kind == FK_ROTATION) {
            convertRotation<true><<<(unsigned)((count + 255) / 256), 256, 0, c->stream>>>(
                c->d_stage, c->d_stageQ, count, surface ? c->d_refL : nullptr, c->d_faceStart, (int)c->B, c->Bpad);
            launchTranspose<true>(c, dev, c->d_stageQ, 4, surface);
            c->launches += 2;
        } else if (kind == FK_STAR || kind == FK_RATE) {
            const size_t n = (size_t)3 * c->nmax * c->Bpad;
            const double g = c->gammaN * c->dtStar;
            launchTranspose<true>(c, c->d_uP, c->d_stage, nc, surface); // padding rows of the scratch are never read back
            if (kind == FK_STAR) axpyCells<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dev, c->d_uP, starPartner(c, dev), -g, n);
            else rebaseStar<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(starPartner(c, dev), dev, c->d_uP, g, n);
            c->launches += 2;
        } else {
            launchTranspose<true>(c, dev, c->d_stage, nc, surface);
            c->launches++;
        }
    } else {
        stageInternal(c, dev, nc, surface, kind, c->d_stage);
        CUDA_OK(c, cudaMemcpyAsync(host, c->d_stage, sizeof(double) * count * nc, cudaMemcpyDeviceToHost, c->stream));
    }
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}
// faceArray: the patch faces of a surface field (kind as for the internal part); otherwise an end array [2][nc][Bpad]
static int xferPatch(bf_ctx* c, double* dev, double* host, int nc, int end, int faceArray, bool toDevice, int kind = FK_PLAIN)
{
    const unsigned grid = (unsigned)((c->B + 255) / 256);
    const size_t bytes = sizeof(double) * c->B * nc;
    const int R = c->nmax + 1;
    if (kind == FK_GAMMA) {
        if (toDevice) return BF_OK;
        gammaField<<<(unsigned)((c->B + 127) / 128), 128, 0, c->stream>>>(makeParams(c), c->d_faceStart, nullptr,
                                                                          end ? nullptr : c->d_stagePatch, end ? c->d_stagePatch : nullptr);
        CUDA_OK(c, cudaMemcpyAsync(host, c->d_stagePatch, bytes, cudaMemcpyDeviceToHost, c->stream));
    } else if (toDevice) {
        CUDA_OK(c, cudaMemcpyAsync(c->d_stagePatch, host, bytes, cudaMemcpyHostToDevice, c->stream));
        if (kind == FK_ROTATION) {
            convertRotation<true><<<grid, 256, 0, c->stream>>>(c->d_stagePatch, c->d_stagePatchQ, c->B, c->d_refL, nullptr, (int)c->B, c->Bpad);
            transposePatch<true><<<grid, 256, 0, c->stream>>>(dev, c->d_stagePatchQ, c->d_nSeg, (int)c->B, c->Bpad, R, 4, end, faceArray);
            c->launches++;
        } else
            transposePatch<true><<<grid, 256, 0, c->stream>>>(dev, c->d_stagePatch, c->d_nSeg, (int)c->B, c->Bpad, R, nc, end, faceArray);
    } else {
        if (kind == FK_ROTATION) {
            transposePatch<false><<<grid, 256, 0, c->stream>>>(dev, c->d_stagePatchQ, c->d_nSeg, (int)c->B, c->Bpad, R, 4, end, faceArray);
            convertRotation<false><<<grid, 256, 0, c->stream>>>(c->d_stagePatch, c->d_stagePatchQ, c->B, c->d_refL, nullptr, (int)c->B, c->Bpad);
            c->launches++;
        } else
            transposePatch<false><<<grid, 256, 0, c->stream>>>(dev, c->d_stagePatch, c->d_nSeg, (int)c->B, c->Bpad, R, nc, end, faceArray);
        CUDA_OK(c, cudaMemcpyAsync(host, c->d_stagePatch, bytes, cudaMemcpyDeviceToHost, c->stream));
    }
    c->launches++;
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}

extern "C" int bf_set_field(bf_ctx* c, int32_t field, const double* internal, const double* left, const double* right)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_field between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    CUDA_OK(c, cudaSetDevice(c->device));
    double *dev, *endArr; int nc, surf, kind;
    int rc = fieldInfo(c, field, &dev, &endArr, &nc, &surf, &kind);
    if (rc) return rc;
    if (internal) {
        if (!dev) return fail(c, BF_ERR_INVALID, "field %d is not allocated in this context", field);
        if ((rc = xferInternal(c, dev, (double*)internal, nc, surf, kind, true))) return rc;
    }
    const double* sides[2] = {left, right};
    for (int end = 0; end < 2; end++) {
        if (!sides[end]) continue;
        if (surf) { if ((rc = xferPatch(c, dev, (double*)sides[end], nc, end, 1, true, kind))) return rc; }
        else if (endArr) {
            if ((rc = xferPatch(c, endArr, (double*)sides[end], nc, end, 0, true))) return rc;
            // a plain fixedValue keeps the value it was given; prevIter == value (CTL.C:1110-1111)
            if (field == BF_FIELD_W && (rc = xferPatch(c, c->d_wPresc, (double*)sides[end], nc, end, 0, true))) return rc;
            if (field == BF_FIELD_THETA && (rc = xferPatch(c, c->d_thPresc, (double*)sides[end], nc, end, 0, true))) return rc;
        }
    }
    return BF_OK;
}

extern "C" int bf_get_field(bf_ctx* c, int32_t field, double* internal, double* left, double* right)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_get_field between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    CUDA_OK(c, cudaSetDevice(c->device));
    double *dev, *endArr; int nc, surf, kind;
    int rc = fieldInfo(c, field, &dev, &endArr, &nc, &surf, &kind);
    if (rc) return rc;
    if (internal) {
        if (!dev) return fail(c, BF_ERR_INVALID, "field %d is not kept in this context (scheme / storeIncrements)", field);
        if ((rc = xferInternal(c, dev, internal, nc, surf, kind, false))) return rc;
    }
    double* sides[2] = {left, right};
    for (int end = 0; end < 2; end++) {
        if (!sides[end]) continue;
        if (surf) { if ((rc = xferPatch(c, dev, sides[end], nc, end, 1, false, kind))) return rc; }
        else if (endArr) { if ((rc = xferPatch(c, endArr, sides[end], nc, end, 0, false))) return rc; }
        else return fail(c, BF_ERR_INVALID, "field %d has no mirrored boundary values", field);
    }
    return BF_OK;
}

// Asynchronous mirror of internal fields into (pinned) host buffers: the fields are transposed
// to the reference ordering into disjoint regions of the staging buffer on the compute stream,
// then copied device->host on a second stream, so the copy of step k overlaps the Newton loop of
// step k+1.  bf_mirror_wait() (or the next bf_mirror_begin) waits for the copies.
extern "C" int bf_mirror_begin(bf_ctx* c, int32_t nFields, const int32_t* fields, double* const* hostInternal)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_mirror_begin between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (nFields < 1 || !fields || !hostInternal) return fail(c, BF_ERR_INVALID, "bf_mirror_begin: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (c->copyPending) { // the staging regions are about to be overwritten
        CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->evCopied, 0));
    }
    size_t off = 0;
    struct Job { double* host; size_t off, count; } jobs[16];
    if (nFields > 16) return fail(c, BF_ERR_INVALID, "bf_mirror_begin: at most 16 fields");
    for (int k = 0; k < nFields; k++) {
        double *dev, *endArr; int nc, surf, kind;
        int rc = fieldInfo(c, fields[k], &dev, &endArr, &nc, &surf, &kind);
        if (rc) return rc;
        if (!dev) return fail(c, BF_ERR_INVALID, "field %d is not kept in this context", fields[k]);
        const int64_t count = surf ? c->F : c->N;
        if (off + (size_t)count * nc > c->stageDoubles) return fail(c, BF_ERR_INVALID, "bf_mirror_begin: fields exceed the staging buffer (9 doubles per cell)");
        stageInternal(c, dev, nc, surf, kind, c->d_stage + off);
        jobs[k] = {hostInternal[k], off, (size_t)count * nc};
        off += (size_t)count * nc;
    }
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaEventRecord(c->evStaged, c->stream));
    CUDA_OK(c, cudaStreamWaitEvent(c->copyStream, c->evStaged, 0));
    for (int k = 0; k < nFields; k++)
        if (jobs[k].count)
            CUDA_OK(c, cudaMemcpyAsync(jobs[k].host, c->d_stage + jobs[k].off, sizeof(double) * jobs[k].count,
                                       cudaMemcpyDeviceToHost, c->copyStream));
    CUDA_OK(c, cudaEventRecord(c->evCopied, c->copyStream));
    c->copyPending = true;
    return BF_OK;
}
extern "C" int bf_mirror_wait(bf_ctx* c)
{
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaStreamSynchronize(c->copyStream));
    c->copyPending = false;
    return BF_OK;
}

extern "C" int bf_set_distributed_loads(bf_ctx* c, const double* q, const double* m)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_distributed_loads between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    CUDA_OK(c, cudaSetDevice(c->device));
    const double* src[2] = {q, m};
    double** dst[2] = {&c->d_q, &c->d_m};
    bool* has[2] = {&c->hasQ, &c->hasM};
    for (int k = 0; k < 2; k++) {
        if (!src[k]) { *has[k] = false; continue; }
        if (!*dst[k]) {
            int rc = devAlloc(c, (void**)dst[k], sizeof(double) * 3 * (size_t)c->nmax * c->Bpad);
            if (rc) return rc;
        }
        int rc = xferInternal(c, *dst[k], (double*)src[k], 3, 0, FK_PLAIN, true);
        if (rc) return rc;
        *has[k] = true;
    }
    return BF_OK;
}

// constant/pointForces (Evolve.C:211-271), evaluated by the host at the current time.  The sweeps need, per cell,
// sum F0 and sum zeta*F0 for each side of the cell (the lever arm is zeta times a per-side vector): 9 doubles.
extern "C" int bf_set_point_forces(bf_ctx* c, int64_t n, const int64_t* cell, const double* zeta, const double* force)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_point_forces between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (n < 0 || (n > 0 && (!cell || !zeta || !force))) return fail(c, BF_ERR_INVALID, "bf_set_point_forces: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (n == 0) { c->hasPf = false; return BF_OK; }
    std::vector<double> h((size_t)c->N * 9, 0.0);
    for (int64_t k = 0; k < n; k++) {
        if (cell[k] < 0 || cell[k] >= c->N) return fail(c, BF_ERR_INVALID, "bf_set_point_forces: cell %lld out of range", (long long)cell[k]);
        double* o = h.data() + 9 * cell[k];
        for (int j = 0; j < 3; j++) {
            o[j] += force[3 * k + j];
            if (zeta[k] > BF_SMALL) o[3 + j] += zeta[k] * force[3 * k + j];
            else if (zeta[k] < -BF_SMALL) o[6 + j] += zeta[k] * force[3 * k + j];
        }
    }
    if (!c->d_pf) {
        int rc = devAlloc(c, (void**)&c->d_pf, sizeof(double) * 9 * (size_t)c->nmax * c->Bpad);
        if (!rc && !c->d_pfSrc) rc = devAlloc(c, (void**)&c->d_pfSrc, sizeof(double) * 6 * (size_t)c->nmax * c->Bpad); // shared with the momentum contributions
        if (rc) return rc;
    }
    int rc = xferInternal(c, c->d_pf, h.data(), 9, 0, FK_PLAIN, true);
    if (rc) return rc;
    c->hasPf = true;
    return BF_OK;
}

// constant/beamMomentumContributionProperties (CTL.C:1137-1182; Evolve.C:539-571): see beam_contrib.cuh
extern "C" int bf_set_momentum_contributions(bf_ctx* c, int32_t n, const bf_momentum_contribution* list, const double* radius,
                                             const double* translation)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_momentum_contributions between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (n < 0 || n > BF_MAX_CONTRIBUTIONS || (n > 0 && (!list || !radius))) return fail(c, BF_ERR_INVALID, "bf_set_momentum_contributions: bad argument");
    for (int i = 0; i < n; i++) {
        if (list[i].type != BF_CONTRIB_MORISON_DRAG && list[i].type != BF_CONTRIB_GROUND_CONTACT)
            return fail(c, BF_ERR_INVALID, "unknown beamMomentumContributionType %d", list[i].type);
        if (list[i].type == BF_CONTRIB_MORISON_DRAG && c->scheme == BF_SCHEME_STEADY) // morisonDragContribution.C:178-190
            return fail(c, BF_ERR_UNSUPPORTED, "Momentum contribution due to Morison drag for time scheme steadyState is not defined!!");
    }
    c->contrib.n = 0;
    if (!n) return BF_OK;
    CUDA_OK(c, cudaSetDevice(c->device));
    const size_t bp = (size_t)c->Bpad;
    if (!c->contrib.radius) {
        double *r = nullptr, *t = nullptr;
        int rc = devAlloc(c, (void**)&r, sizeof(double) * bp);
        if (!rc) rc = devAlloc(c, (void**)&t, sizeof(double) * 3 * bp);
        if (!rc && !c->d_pfSrc) rc = devAlloc(c, (void**)&c->d_pfSrc, sizeof(double) * 6 * (size_t)c->nmax * bp);
        if (rc) return rc;
        c->contrib.radius = r; c->contrib.translation = t;
    }
    std::vector<double> h(4 * bp, 0.0);
    for (int64_t b = 0; b < c->B; b++) {
        h[b] = radius[b];
        if (translation) for (int k = 0; k < 3; k++) h[(1 + k) * bp + b] = translation[3 * b + k];
    }
    CUDA_OK(c, cudaMemcpyAsync((void*)c->contrib.radius, h.data(), sizeof(double) * bp, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync((void*)c->contrib.translation, h.data() + bp, sizeof(double) * 3 * bp, cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; i++) {
        c->contrib.type[i] = list[i].type;
        for (int k = 0; k < 6; k++) c->contrib.coeffs[i][k] = list[i].coeffs[k];
    }
    c->contrib.n = n;
    return BF_OK;
}

// One host->device copy of a [B][3] patch array into the staging buffer, then any number of scatters into end arrays.
static int stagePatchUpload(bf_ctx* c, const double* host)
{
    CUDA_OK(c, cudaMemcpyAsync(c->d_stagePatch, host, sizeof(double) * c->B * 3, cudaMemcpyHostToDevice, c->stream));
    return BF_OK;
}
static void scatterPatch(bf_ctx* c, double* endArr, int end)
{
    transposePatch<true><<<(unsigned)((c->B + 255) / 256), 256, 0, c->stream>>>(endArr, c->d_stagePatch, c->d_nSeg, (int)c->B, c->Bpad,
                                                                                c->nmax + 1, 3, end, 0);
    c->launches++;
}
static int setBcValues(bf_ctx* c, int32_t end, int32_t kind, const double* v, bool wait);
extern "C" int bf_set_bc_values(bf_ctx* c, int32_t end, int32_t kind, const double* v) { return setBcValues(c, end, kind, v, true); }
extern "C" int bf_set_bc_values_async(bf_ctx* c, int32_t end, int32_t kind, const double* v) { return setBcValues(c, end, kind, v, false); }
static int setBcValues(bf_ctx* c, int32_t end, int32_t kind, const double* v, bool wait)
{
    if ((end != BF_END_LEFT && end != BF_END_RIGHT) || !v) return fail(c, BF_ERR_INVALID, "bf_set_bc_values: bad argument");
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_bc_values between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    CUDA_OK(c, cudaSetDevice(c->device));
    // entries of beams whose BC class does not use `kind` are never read by the kernels.  One copy of the host array, one
    // wait at the end (the caller may reuse its buffer on return); the scatters stay asynchronous on the stream.
    double* dst = nullptr;
    switch (kind) {
    case BF_BCV_W: dst = c->d_wPresc; break;
    case BF_BCV_THETA: dst = c->d_thPresc; break;
    case BF_BCV_MOMENT: dst = c->d_moment; break;
    case BF_BCV_FORCE: break;
    default: return fail(c, BF_ERR_INVALID, "bf_set_bc_values: unknown kind %d", kind);
    }
    if (dst) {
        if (int rc = stagePatchUpload(c, v)) return rc;
        scatterPatch(c, dst, end);
    } else {
        // forceSeries -> force(), followerForceSeries -> followerForce_, offsetForceSeries -> offsetForce_.
        // force() of a spring patch is its own state (set in evaluate) and must not be overwritten.
        const int32_t* wt = c->hostWType.data() + (size_t)end * c->B;
        bool anySpring = false;
        for (int64_t b = 0; b < c->B; b++) if (wt[b] == BF_W_SPRING) { anySpring = true; break; }
        if (int rc = stagePatchUpload(c, v)) return rc;
        scatterPatch(c, c->d_follower, end);
        scatterPatch(c, c->d_offset, end);
        if (!anySpring) scatterPatch(c, c->d_force, end);
        else {
            std::vector<double> cur((size_t)c->B * 3), in(v, v + (size_t)c->B * 3);
            if (int rc = xferPatch(c, c->d_force, cur.data(), 3, end, 0, false)) return rc;
            for (int64_t b = 0; b < c->B; b++)
                if (wt[b] == BF_W_SPRING) for (int k = 0; k < 3; k++) in[3 * b + k] = cur[3 * b + k];
            return xferPatch(c, c->d_force, in.data(), 3, end, 0, true);
        }
    }
    CUDA_OK(c, cudaGetLastError());
    if (wait) CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}

// Patch values of W and Theta of one end of every beam in ONE call (what functionObjects/beamDisplacements reads every
// step): two gathers, two copies, one wait.
extern "C" int bf_get_patch_pair(bf_ctx* c, int32_t end, double* W, double* Theta)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_get_patch_pair between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if ((end != BF_END_LEFT && end != BF_END_RIGHT) || !W || !Theta) return fail(c, BF_ERR_INVALID, "bf_get_patch_pair: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    const unsigned grid = (unsigned)((c->B + 255) / 256);
    const size_t n = (size_t)c->B * 3;
    transposePatch<false><<<grid, 256, 0, c->stream>>>(c->d_Wb, c->d_stagePatch, c->d_nSeg, (int)c->B, c->Bpad, c->nmax + 1, 3, end, 0);
    transposePatch<false><<<grid, 256, 0, c->stream>>>(c->d_Thb, c->d_stagePatch + n, c->d_nSeg, (int)c->B, c->Bpad, c->nmax + 1, 3, end, 0);
    c->launches += 2;
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaMemcpyAsync(W, c->d_stagePatch, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(Theta, c->d_stagePatch + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}

static int stagePatchPair(bf_ctx* c, int end, double* W, double* Theta)
{
    const unsigned grid = (unsigned)((c->B + 255) / 256);
    const size_t n = (size_t)c->B * 3;
    transposePatch<false><<<grid, 256, 0, c->stream>>>(c->d_Wb, c->d_stageOut, c->d_nSeg, (int)c->B, c->Bpad, c->nmax + 1, 3, end, 0);
    transposePatch<false><<<grid, 256, 0, c->stream>>>(c->d_Thb, c->d_stageOut + n, c->d_nSeg, (int)c->B, c->Bpad, c->nmax + 1, 3, end, 0);
    c->launches += 2;
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaMemcpyAsync(W, c->d_stageOut, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(Theta, c->d_stageOut + n, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    return BF_OK;
}
// Step outputs: patch W and Theta of one end of every beam, delivered by bf_evolve itself (see beamgpu.h)
extern "C" int bf_set_step_outputs(bf_ctx* c, int32_t end, double* W, double* Theta)
{
    if ((end != BF_END_LEFT && end != BF_END_RIGHT) || (!W) != (!Theta)) return fail(c, BF_ERR_INVALID, "bf_set_step_outputs: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (W && !c->d_stageOut && devAlloc(c, (void**)&c->d_stageOut, sizeof(double) * (size_t)c->B * 6)) return fail(c, BF_ERR_NOMEM, "%s", c->err);
    c->outEnd = end; c->outW = W; c->outTheta = Theta;
    return BF_OK;
}

extern "C" int bf_begin_step(bf_ctx* c, double deltaT)
{
    if (!(deltaT > 0)) return fail(c, BF_ERR_INVALID, "bf_begin_step: deltaT must be > 0");
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_begin_step between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    c->deltaT = deltaT;
    c->iOuterCorr = 0;
    c->timeAdvanced = true;
    return BF_OK;
}

extern "C" int bf_set_timing(bf_ctx* c, int32_t enabled) { c->timing = enabled; return BF_OK; }
// A pair of CUDA events on the context's stream bracketing a region chosen by the caller
// (bench.py: the K timed steps).  which = 0 start, 1 stop.
extern "C" int bf_region_mark(bf_ctx* c, int32_t which)
{
    if (which < 0 || which > 1) return fail(c, BF_ERR_INVALID, "bf_region_mark: which must be 0 or 1");
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventRecord(c->evRegion[which], c->stream));
    return BF_OK;
}
extern "C" int bf_region_elapsed_ms(bf_ctx* c, double* ms)
{
    float t = 0;
    CUDA_OK(c, cudaSetDevice(c->device));
    CUDA_OK(c, cudaEventSynchronize(c->evRegion[1]));
    CUDA_OK(c, cudaEventElapsedTime(&t, c->evRegion[0], c->evRegion[1]));
    *ms = t;
    return BF_OK;
}
// functionObjects/beamEnergyData: per-beam energies from the device state (no field leaves the device)
extern "C" int bf_beam_energy(bf_ctx* c, double* out)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_beam_energy between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (!out) return fail(c, BF_ERR_INVALID, "bf_beam_energy: null argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (int rc = drainMirror(c)) return rc;
    if ((size_t)c->B * 3 > c->stageDoubles) return fail(c, BF_ERR_INVALID, "bf_beam_energy: staging buffer too small");
    const BeamParams p = makeParams(c); // gdtPrev = gamma * (the step the Newmark constants belong to)
    const unsigned grid = (unsigned)((c->B + 127) / 128);
    if (c->scheme == BF_SCHEME_NEWMARK) beam_energy<BFK_NEWMARK><<<grid, 128, 0, c->stream>>>(p, c->d_stage);
    else if (c->scheme == BF_SCHEME_EULER) beam_energy<BFK_EULER><<<grid, 128, 0, c->stream>>>(p, c->d_stage);
    else beam_energy<BFK_STEADY><<<grid, 128, 0, c->stream>>>(p, c->d_stage);
    c->launches++;
    CUDA_OK(c, cudaGetLastError());
    CUDA_OK(c, cudaMemcpyAsync(out, c->d_stage, sizeof(double) * 3 * (size_t)c->B, cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    return BF_OK;
}

static void readChunkTimes(bf_ctx* c);
extern "C" int bf_kernel_time_ms(bf_ctx* c, int32_t reset, double* fwd, double* bwd, int64_t* it)
{
    readChunkTimes(c);
    if (fwd) *fwd = c->fwdMs;
    if (bwd) *bwd = c->bwdMs;
    if (it) *it = c->timedIters;
    if (reset) { c->fwdMs = c->bwdMs = 0; c->timedIters = 0; }
    return BF_OK;
}

static BeamParams makeParams(bf_ctx* c)
{
    BeamParams p;
    memset(&p, 0, sizeof p);
    p.B = (int)c->B; p.Bpad = c->Bpad; p.nmax = c->nmax; p.hL = c->hL;
    p.newmark = c->scheme; p.first = c->timeAdvanced ? 1 : 0; p.storeInc = c->storeInc;
    p.dt = c->deltaT; p.beta = c->betaN; p.gamma = c->gammaN;
    p.nk1 = 1.0 / (p.dt * p.beta); p.nk2 = 0.5 / p.beta - 1.0; p.nVel = p.gamma / (p.beta * p.dt);
    p.nAcc = 1.0 / (p.beta * p.dt * p.dt); p.nU = (1.0 / p.dt) * (p.gamma / p.beta);
    p.gdt = p.gamma * p.dt; p.gdtPrev = p.gamma * c->dtStar;
    p.nSeg = c->d_nSeg; p.dZ = c->d_dZ; p.CQ = c->d_CQ; p.CM = c->d_CM; p.ARho = c->d_ARho; p.CI = c->d_CI;
    p.weight = c->hasWeight ? c->d_weight : nullptr; p.refL = c->d_refL; p.wType = c->d_wType; p.thType = c->d_thType;
    p.W = c->d_W; p.Th = c->d_Th; p.Lam = c->d_Lam; p.U = c->d_U; p.Ac = c->d_Ac; p.Om = c->d_Om; p.dOm = c->d_dOm;
    p.DW = c->d_DW; p.DTh = c->d_DTh;
    p.q = c->hasQ ? c->d_q : nullptr; p.m = c->hasM ? c->d_m : nullptr; p.Lc = c->d_Lc;
    p.pf = c->hasPf ? c->d_pf : nullptr; p.pfSrc = (c->hasPf || c->contrib.n) ? c->d_pfSrc : nullptr;
    p.Wold = c->d_Wold; p.Uold = c->d_Uold; p.Omold = c->d_Omold; p.Lamold = c->d_Lamold;
    p.Lf = c->d_Lf; p.K = c->d_K; p.Q = c->d_Q; p.M = c->d_M;
    p.Wb = c->d_Wb; p.Thb = c->d_Thb; p.Lamb = c->d_Lamb; p.DWb = c->d_DWb; p.DThb = c->d_DThb; p.force = c->d_force;
    p.wPresc = c->d_wPresc; p.thPresc = c->d_thPresc; p.follower = c->d_follower; p.offset = c->d_offset;
    p.moment = c->d_moment; p.springK = c->d_springK; p.springDir = c->d_springDir; p.ifW = c->d_ifW; p.wSave = c->d_Wsave;
    p.uP = c->d_uP; p.bP = c->d_bP; p.partial = c->d_partial;
    p.nTiles = c->partialStride;
    p.cv = c->cvNow;
    return p;
}

#define LAUNCH_SCHEME_T(KERNEL, GRID, THR, SMEM, P)                                                   \
    do {                                                                                              \
        if (c->scheme == BF_SCHEME_NEWMARK) KERNEL<BFK_NEWMARK><<<(GRID), (THR), (SMEM), c->stream>>>(P); \
        else if (c->scheme == BF_SCHEME_EULER) KERNEL<BFK_EULER><<<(GRID), (THR), (SMEM), c->stream>>>(P); \
        else KERNEL<BFK_STEADY><<<(GRID), (THR), (SMEM), c->stream>>>(P);                             \
    } while (0)
#define LAUNCH_SCHEME(KERNEL, GRID, SMEM, P)                                                          \
    do {                                                                                              \
        if (c->scheme == BF_SCHEME_NEWMARK) KERNEL<BFK_NEWMARK><<<(GRID), THREADS, (SMEM), c->stream>>>(P); \
        else if (c->scheme == BF_SCHEME_EULER) KERNEL<BFK_EULER><<<(GRID), THREADS, (SMEM), c->stream>>>(P); \
        else KERNEL<BFK_STEADY><<<(GRID), THREADS, (SMEM), c->stream>>>(P);                           \
    } while (0)
// The warp-specialised forward kernel (beam_coop.cuh): 16 or 32 beams per block, whole beams or halves (grid.y = 2).
template <int BPB, bool SPLIT, int EW>
static void launchForwardWsT(bf_ctx* c, const BeamParams& p)
{
    const dim3 g((unsigned)((c->Bpad + BPB - 1) / BPB), SPLIT ? 2 : 1);
    if (c->scheme == BF_SCHEME_NEWMARK) beam_forward_ws<BFK_NEWMARK, BPB, SPLIT, EW><<<g, WS_THREADS(EW), WS_SMEM_BYTES(BPB, EW), c->stream>>>(p);
    else if (c->scheme == BF_SCHEME_EULER) beam_forward_ws<BFK_EULER, BPB, SPLIT, EW><<<g, WS_THREADS(EW), WS_SMEM_BYTES(BPB, EW), c->stream>>>(p);
    else beam_forward_ws<BFK_STEADY, BPB, SPLIT, EW><<<g, WS_THREADS(EW), WS_SMEM_BYTES(BPB, EW), c->stream>>>(p);
}
static int launchForwardWs(bf_ctx* c, const BeamParams& p) // returns the number of blocks (partial sums of slot 0)
{
    if (c->wsBpb == 32 && c->wsEw == 2) { if (c->split) launchForwardWsT<32, true, 2>(c, p); else launchForwardWsT<32, false, 2>(c, p); }
    else if (c->wsBpb == 32) { if (c->split) launchForwardWsT<32, true, 1>(c, p); else launchForwardWsT<32, false, 1>(c, p); }
    else { if (c->split) launchForwardWsT<16, true, 1>(c, p); else launchForwardWsT<16, false, 1>(c, p); }
    return ((c->Bpad + c->wsBpb - 1) / c->wsBpb) * (c->split ? 2 : 1);
}
// beam_update: one thread per cell; one warp per block (updateWarpBlocks) or four
static int updateGrid(const bf_ctx* c)
{
    const size_t cellsPadded = (size_t)c->Bpad * (size_t)c->nmax;
    return c->updateWarpBlocks ? (int)((cellsPadded + 31) / 32) : (int)((cellsPadded + 127) / 128);
}
static void launchUpdate(bf_ctx* c, const BeamParams& p, int cellGrid)
{
    if (c->updateWarpBlocks) {
        if (c->scheme == BF_SCHEME_NEWMARK) beam_update<BFK_NEWMARK, true><<<cellGrid, 32, 0, c->stream>>>(p, c->d_X);
        else if (c->scheme == BF_SCHEME_EULER) beam_update<BFK_EULER, true><<<cellGrid, 32, 0, c->stream>>>(p, c->d_X);
        else beam_update<BFK_STEADY, true><<<cellGrid, 32, 0, c->stream>>>(p, c->d_X);
    } else {
        if (c->scheme == BF_SCHEME_NEWMARK) beam_update<BFK_NEWMARK, false><<<cellGrid, 128, 0, c->stream>>>(p, c->d_X);
        else if (c->scheme == BF_SCHEME_EULER) beam_update<BFK_EULER, false><<<cellGrid, 128, 0, c->stream>>>(p, c->d_X);
        else beam_update<BFK_STEADY, false><<<cellGrid, 128, 0, c->stream>>>(p, c->d_X);
    }
}

// Launches one Newton iteration; the three LOCAL sums of squares end up in d_sums.
// evs: the three timing events of this iteration (bf_evolve's pipelined loop passes one set per enqueued iteration)
static int launchIteration(bf_ctx* c, cudaEvent_t* evs = nullptr)
{
    if (!evs) evs = c->ev;
    BeamParams p = makeParams(c);
    // pipelined loop: the reduction of this iteration will run beside the forward part of the next one, which writes its
    // partial sums into the other bank
    const bool sideReduce = p.cv.xbuf && c->sideReduce;
    if (sideReduce) p.partial = c->d_partial + (size_t)(p.cv.iter & 1) * 3 * (size_t)c->partialStride;
    // between the forward and the backward part: the backward part starts with the verdict of the previous iteration
#define MID_ITERATION()                                                                                     \
    do {                                                                                                    \
        if (c->timing) CUDA_OK(c, cudaEventRecord(evs[1], c->stream));                                      \
        if (c->redPending) { CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->evReduced, 0)); c->redPending = false; } \
    } while (0)
    if (p.pf) { // lever arms follow the current W: re-evaluated every iteration (Evolve.C:221)
        beam_point_forces<<<(unsigned)((c->B + 127) / 128), 128, 0, c->stream>>>(p);
        c->launches++;
    }
    if (c->contrib.n) { // beamMomentumContribution sources follow the current W and U: every iteration (Evolve.C:539-571)
        ContribParams cp = c->contrib;
        cp.keepPointForces = p.pf != nullptr;
        const unsigned cgrid = (unsigned)(((size_t)((c->B + 31) / 32) * 32 * (size_t)c->nmax + 127) / 128);
        if (c->scheme == BF_SCHEME_NEWMARK) beam_momentum_contributions<BFK_NEWMARK><<<cgrid, 128, 0, c->stream>>>(p, cp);
        else if (c->scheme == BF_SCHEME_EULER) beam_momentum_contributions<BFK_EULER><<<cgrid, 128, 0, c->stream>>>(p, cp);
        else beam_momentum_contributions<BFK_STEADY><<<cgrid, 128, 0, c->stream>>>(p, cp);
        c->launches++;
    }
    if (c->timing) CUDA_OK(c, cudaEventRecord(evs[0], c->stream));
    int nF = c->grid, nB = c->grid; // blocks that wrote the partial sums of the forward / backward part
    if (c->nPairs) { // coupled bundle: cell-parallel assembly, couplings, BiCGStab preconditioned by the per-beam solve, cell-parallel update
        LongParams q;
        memset(&q, 0, sizeof q);
        q.N = c->N; q.cellBeam = c->d_cellBeam; q.cellStart = c->d_cellStart; q.X = c->d_longX;
        q.L0 = c->d_longRows[8]; q.D0 = c->d_longRows[9]; q.U0 = c->d_longRows[10]; q.R0 = c->d_longRows[11];
        CoupledParams cq;
        cq.N = c->N; cq.B = (int)c->B; cq.cellStart = c->d_cellStart; cq.cellBeam = c->d_cellBeam;
        cq.adjPtr = c->d_adjPtr; cq.adjCol = c->d_adjCol; cq.adjK = c->d_adjK;
        cq.L = q.L0; cq.D = q.D0; cq.U = q.U0; cq.R = q.R0; cq.Dinv = c->d_cplDinv; cq.UP = c->d_cplUP;
        const int gb = (int)((c->N + 127) / 128);
        if (c->scheme == BF_SCHEME_NEWMARK) long_assemble<BFK_NEWMARK><<<gb, 128, 0, c->stream>>>(p, q);
        else if (c->scheme == BF_SCHEME_EULER) long_assemble<BFK_EULER><<<gb, 128, 0, c->stream>>>(p, q);
        else long_assemble<BFK_STEADY><<<gb, 128, 0, c->stream>>>(p, q);
        cpl_add_couplings<<<gb, 128, 0, c->stream>>>(p, cq);
        c->launches += 2;
        if (int rc = solveCoupled(c, cq)) return rc;
        // ||source||^2 of the coupled system (the assembly's partial sums do not contain the spring forces)
        cpl_dots<<<1, 256, 0, c->stream>>>(6 * c->N, cq.R, cq.R, nullptr, nullptr, nullptr, nullptr, nullptr, c->d_hlin);
        cpl_set_partial0<<<1, 256, 0, c->stream>>>(c->d_partial, gb, c->d_hlin);
        if (c->timing) CUDA_OK(c, cudaEventRecord(evs[1], c->stream));
        long_update_faces<<<gb, 128, 0, c->stream>>>(p, q);
        if (c->scheme == BF_SCHEME_NEWMARK) long_update_cells<BFK_NEWMARK><<<gb, 128, 0, c->stream>>>(p, q);
        else if (c->scheme == BF_SCHEME_EULER) long_update_cells<BFK_EULER><<<gb, 128, 0, c->stream>>>(p, q);
        else long_update_cells<BFK_STEADY><<<gb, 128, 0, c->stream>>>(p, q);
        c->launches += 4;
        nF = nB = gb;
    } else if (c->longMode) {
        nF = nB = c->longBlocks; // one thread per cell: assemble, block cyclic reduction, update
        LongParams q;
        q.N = c->N; q.cellBeam = c->d_cellBeam; q.cellStart = c->d_cellStart; q.X = c->d_longX;
        for (int k = 0; k < 2; k++) { q.L[k] = c->d_longRows[4 * k]; q.D[k] = c->d_longRows[4 * k + 1]; q.U[k] = c->d_longRows[4 * k + 2]; q.R[k] = c->d_longRows[4 * k + 3]; }
        q.L0 = c->d_longRows[8]; q.D0 = c->d_longRows[9]; q.U0 = c->d_longRows[10]; q.R0 = c->d_longRows[11];
        q.src = 0; q.step = 0; q.refine = 0;
        const int gb = c->longBlocks;
        if (c->scheme == BF_SCHEME_NEWMARK) long_assemble<BFK_NEWMARK><<<gb, 128, 0, c->stream>>>(p, q);
        else if (c->scheme == BF_SCHEME_EULER) long_assemble<BFK_EULER><<<gb, 128, 0, c->stream>>>(p, q);
        else long_assemble<BFK_STEADY><<<gb, 128, 0, c->stream>>>(p, q);
        c->launches++;
        for (int pass = 0; pass < 2; pass++) { // solve, then one step of iterative refinement on the residual
            q.refine = pass; q.src = 0;
            long_pcr_load<<<gb, 128, 0, c->stream>>>(p, q);
            for (int s = 1; s < c->nmax; s *= 2) {
                q.step = s;
                long_pcr_step<<<(unsigned)((c->N * PCR_COLS + 63) / 64), 64, 0, c->stream>>>(p, q);
                q.src = 1 - q.src;
                c->launches++;
            }
            long_pcr_finish<<<gb, 128, 0, c->stream>>>(q);
            c->launches += 2;
        }
        if (c->timing) CUDA_OK(c, cudaEventRecord(evs[1], c->stream));
        long_update_faces<<<gb, 128, 0, c->stream>>>(p, q);
        if (c->scheme == BF_SCHEME_NEWMARK) long_update_cells<BFK_NEWMARK><<<gb, 128, 0, c->stream>>>(p, q);
        else if (c->scheme == BF_SCHEME_EULER) long_update_cells<BFK_EULER><<<gb, 128, 0, c->stream>>>(p, q);
        else long_update_cells<BFK_STEADY><<<gb, 128, 0, c->stream>>>(p, q);
        c->launches += 2;
    } else if (c->split) { // two threads per beam, the halves meet in the middle
        const dim3 g2((unsigned)c->grid, 2);
        nF = nB = 2 * c->grid;
        if (c->coopFwd) nF = launchForwardWs(c, p); // assembler + eliminator warps per half beam
        else if (c->coopBwd) { // the cell-parallel update reads the pre-update W the sweep saves
            if (c->scheme == BF_SCHEME_NEWMARK) beam_forward_split<BFK_NEWMARK, true><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
            else if (c->scheme == BF_SCHEME_EULER) beam_forward_split<BFK_EULER, true><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
            else beam_forward_split<BFK_STEADY, true><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
        } else {
            if (c->scheme == BF_SCHEME_NEWMARK) beam_forward_split<BFK_NEWMARK, false><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
            else if (c->scheme == BF_SCHEME_EULER) beam_forward_split<BFK_EULER, false><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
            else beam_forward_split<BFK_STEADY, false><<<g2, THREADS, SMEM_BYTES, c->stream>>>(p);
        }
        MID_ITERATION();
        if (c->coopBwd) { // small bundles: serial back-substitution of the two halves, then the update with one thread per cell
            const int cellGrid = updateGrid(c);
            const dim3 gs((unsigned)(c->Bpad / 32), 2);
            beam_backsub_split<<<gs, 32, BSS_SMEM_BYTES, c->stream>>>(p, c->d_X);
            launchUpdate(c, p, cellGrid);
            c->launches += 1;
            nB = cellGrid;
        } else
            LAUNCH_SCHEME(beam_backward_split, g2, 0, p);
        c->launches += 2;
    } else if (c->splitKernels || c->coopFwd || c->coopBwd) { // product path: one launch per sweep
        int nWs = 0;
        if (c->coopFwd) nWs = launchForwardWs(c, p);
        else LAUNCH_SCHEME(beam_forward, c->grid, SMEM_BYTES, p);
        MID_ITERATION();
        const int cellGrid = updateGrid(c);
        if (c->coopBwd) { // serial back-substitution, then the update with one thread per cell
            if (!c->coopFwd) // beam_forward does not save the pre-update W (a combination only tests force): a device copy of the array
                CUDA_OK(c, cudaMemcpyAsync(c->d_Wsave, c->d_W, sizeof(double) * 3 * (size_t)c->nmax * (size_t)c->Bpad, cudaMemcpyDeviceToDevice, c->stream));
            beam_backsub<<<(unsigned)((c->Bpad + BS_THREADS - 1) / BS_THREADS), BS_THREADS, BS_SMEM_BYTES, c->stream>>>(p, c->d_X);
            launchUpdate(c, p, cellGrid);
            c->launches += 1;
        } else LAUNCH_SCHEME(beam_backward, c->grid, 0, p);
        c->launches += 2;
        nF = c->coopFwd ? nWs : c->grid;
        nB = c->coopBwd ? cellGrid : c->grid;
    } else {
        LAUNCH_SCHEME(beam_newton, c->grid, SMEM_BYTES, p);
        if (c->timing) CUDA_OK(c, cudaEventRecord(evs[1], c->stream));
        c->launches += 1;
    }
    if (c->timing) CUDA_OK(c, cudaEventRecord(evs[2], c->stream));
    if (p.cv.xbuf) { // pipelined loop: the sums go into every rank's exchange buffer (one-sided), nobody waits here
        PeerTable pt;
        const size_t bank = (size_t)(c->epoch & 1) * CONV_MAXK * c->xRanks * 4;
        for (int r = 0; r < CONV_MAX_RANKS; r++) pt.bank[r] = r < c->xRanks ? c->peerXbuf[r] + bank : nullptr;
        if (sideReduce) {
            CUDA_OK(c, cudaEventRecord(c->evSwept, c->stream));
            CUDA_OK(c, cudaStreamWaitEvent(c->redStream, c->evSwept, 0));
            beam_reduce_exchange<<<1, RED_THREADS, 0, c->redStream>>>(p.partial, c->partialStride, nF, nB, p.cv, pt, c->xRank);
            CUDA_OK(c, cudaEventRecord(c->evReduced, c->redStream));
            c->redPending = true;
        } else
            beam_reduce_exchange<<<1, RED_THREADS, 0, c->stream>>>(p.partial, c->partialStride, nF, nB, p.cv, pt, c->xRank);
    } else
        beam_reduce_partials<<<1, RED_THREADS, 0, c->stream>>>(c->d_partial, c->partialStride, nF, nB, c->d_sums, c->comm ? nullptr : c->d_hsums);
    c->launches += 1;
    CUDA_OK(c, cudaGetLastError());
    if (p.first) c->dtStar = c->deltaT; // the backward sweep of a first iteration re-bases U*, Omega* on this step
    c->timeAdvanced = false;
    c->iOuterCorr++;
    return BF_OK;
}

static int fetchSums(bf_ctx* c, double sumsq[3])
{
    if (c->comm) { // the all-reduced sums live in device memory: publish them to the mapped host words
        beam_publish_sums<<<1, 32, 0, c->stream>>>(c->d_sums, c->d_hsums);
        c->launches++;
    }
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; k++) sumsq[k] = c->h_sums[k];
    if (c->timing) {
        float a = 0, b = 0;
        CUDA_OK(c, cudaEventElapsedTime(&a, c->ev[0], c->ev[1]));
        CUDA_OK(c, cudaEventElapsedTime(&b, c->ev[1], c->ev[2]));
        c->fwdMs += a; c->bwdMs += b; c->timedIters++;
    }
    return BF_OK;
}

extern "C" int bf_newton_iteration(bf_ctx* c, double sumsq[3])
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_newton_iteration between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    CUDA_OK(c, cudaSetDevice(c->device));
    int rc = launchIteration(c);
    if (rc) return rc;
    return fetchSums(c, sumsq);
}

// ---- coupled bundles: bf_set_couplings, the BiCGStab loop (myBlockBiCGStabSolver.C:79-189) ---------------------------------
extern "C" int bf_set_couplings(bf_ctx* c, int64_t nPairs, const int64_t* cellA, const int64_t* cellB, const double* stiffness)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_set_couplings between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (nPairs < 0 || (nPairs > 0 && (!cellA || !cellB || !stiffness))) return fail(c, BF_ERR_INVALID, "bf_set_couplings: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    c->nPairs = 0;
    if (!nPairs) return BF_OK;
    const size_t N = (size_t)c->N;
    std::vector<int64_t> ptr(N + 1, 0), col(2 * (size_t)nPairs), fill(N, 0);
    std::vector<double> kv(2 * (size_t)nPairs);
    for (int64_t p = 0; p < nPairs; p++) {
        if (cellA[p] < 0 || cellA[p] >= c->N || cellB[p] < 0 || cellB[p] >= c->N || cellA[p] == cellB[p])
            return fail(c, BF_ERR_INVALID, "bf_set_couplings: pair %lld out of range", (long long)p);
        ptr[cellA[p] + 1]++; ptr[cellB[p] + 1]++;
    }
    for (size_t g = 0; g < N; g++) ptr[g + 1] += ptr[g];
    for (int64_t p = 0; p < nPairs; p++) { // in pair order: the gather of a cell adds its partners in the order the host gave them
        const int64_t a = cellA[p], b = cellB[p];
        col[ptr[a] + fill[a]] = b; kv[ptr[a] + fill[a]++] = stiffness[p];
        col[ptr[b] + fill[b]] = a; kv[ptr[b] + fill[b]++] = stiffness[p];
    }
    if (int rc = ensureLongBuffers(c)) return rc;
    if (!c->d_cplVec) {
        int rc = devAlloc(c, (void**)&c->d_cplDinv, sizeof(double) * 36 * N);
        if (!rc) rc = devAlloc(c, (void**)&c->d_cplUP, sizeof(double) * 36 * N);
        if (!rc) rc = devAlloc(c, (void**)&c->d_cplVec, sizeof(double) * 8 * 6 * N);
        if (!rc) rc = devAlloc(c, (void**)&c->d_adjPtr, sizeof(int64_t) * (N + 1));
        if (rc) return rc;
        CUDA_OK(c, cudaHostAlloc((void**)&c->h_lin, 16 * sizeof(double), cudaHostAllocMapped));
        CUDA_OK(c, cudaHostGetDevicePointer((void**)&c->d_hlin, c->h_lin, 0));
    }
    // the adjacency arrays are re-allocated per call (the pair count may change); the pool frees them with the context
    if (int rc = devAlloc(c, (void**)&c->d_adjCol, sizeof(int64_t) * col.size())) return rc;
    if (int rc = devAlloc(c, (void**)&c->d_adjK, sizeof(double) * kv.size())) return rc;
    CUDA_OK(c, cudaMemcpyAsync(c->d_adjPtr, ptr.data(), sizeof(int64_t) * (N + 1), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_adjCol, col.data(), sizeof(int64_t) * col.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(c->d_adjK, kv.data(), sizeof(double) * kv.size(), cudaMemcpyHostToDevice, c->stream));
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    c->nPairs = nPairs;
    return BF_OK;
}
extern "C" int bf_set_linear_solver(bf_ctx* c, int32_t maxIter, double tolerance, double relTol)
{
    if (maxIter < 1 || !(tolerance >= 0) || !(relTol >= 0)) return fail(c, BF_ERR_INVALID, "bf_set_linear_solver: bad argument");
    c->linMaxIter = maxIter; c->linTol = tolerance; c->linRelTol = relTol;
    return BF_OK;
}
extern "C" int32_t bf_linear_iterations(const bf_ctx* c) { return c->linIterations; }

// Solves (T + C) x = R for the rows assembled in q; x = d_longX.  The loop is the reference's, statement by statement;
// the scalars live on the host (three synchronisations per iteration: this path is not the tuned one).
static int solveCoupled(bf_ctx* c, const CoupledParams& q)
{
    const int64_t n = 6 * c->N;
    double *r = c->d_cplVec, *rw = r + n, *pv = r + 2 * n, *ph = r + 3 * n, *v = r + 4 * n, *sv = r + 5 * n, *sh = r + 6 * n, *t = r + 7 * n;
    double* x = c->d_longX;
    const unsigned gv = (unsigned)((n + 255) / 256), gc = (unsigned)((c->N + 127) / 128), gb = (unsigned)((c->B + 63) / 64);
    auto dots = [&](const double* a0, const double* b0, const double* a1, const double* b1, const double* a2, const double* b2, const double* l1) {
        cpl_dots<<<1, 256, 0, c->stream>>>(n, a0, b0, a1, b1, a2, b2, l1, c->d_hlin);
        c->launches++;
        return cudaStreamSynchronize(c->stream);
    };
    cpl_factor<<<gb, 64, 0, c->stream>>>(q);
    CUDA_OK(c, cudaMemsetAsync(x, 0, sizeof(double) * n, c->stream));
    CUDA_OK(c, cudaMemcpyAsync(r, q.R, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream)); // r = b - A x0, x0 = 0
    CUDA_OK(c, cudaMemcpyAsync(rw, q.R, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
    CUDA_OK(c, cudaMemsetAsync(pv, 0, sizeof(double) * n, c->stream));
    CUDA_OK(c, cudaMemsetAsync(v, 0, sizeof(double) * n, c->stream));
    c->launches++;
    CUDA_OK(c, dots(rw, r, nullptr, nullptr, nullptr, nullptr, r));
    double norm[6], rho = 1e20, rhoNext = c->h_lin[0];
    for (int k = 0; k < 6; k++) norm[k] = c->h_lin[3 + k] + 1e-20; // normFactor with x0 = 0: gSum(cmptMag(b)) + small
    auto measure = [&]() { double m = 0; for (int k = 0; k < 6; k++) if (c->h_lin[3 + k] / norm[k] > m) m = c->h_lin[3 + k] / norm[k]; return m; };
    const double initial = measure();
    double final = initial;
    c->linIterations = 0;
    auto stop = [&]() { return c->linIterations >= c->linMaxIter || final < c->linTol || (c->linRelTol > 1e-20 && final <= c->linRelTol * initial); };
    if (!stop()) {
        double rhoOld, alpha = 0, omega = 1e20, beta;
        do {
            rhoOld = rho;
            rho = rhoNext; // gSumProd(rw, r)
            beta = rho / rhoOld * (alpha / omega);
            if (rho == 0) { // restart if breakdown occurs
                CUDA_OK(c, cudaMemcpyAsync(rw, r, sizeof(double) * n, cudaMemcpyDeviceToDevice, c->stream));
                CUDA_OK(c, dots(rw, r, nullptr, nullptr, nullptr, nullptr, nullptr));
                rho = c->h_lin[0];
                alpha = 0; omega = 0; beta = 0;
            }
            cpl_update_p<<<gv, 256, 0, c->stream>>>(n, pv, r, v, beta, omega);
            cpl_precond<<<gb, 64, 0, c->stream>>>(q, pv, ph);
            cpl_amul<<<gc, 128, 0, c->stream>>>(q, ph, v);
            CUDA_OK(c, dots(rw, v, nullptr, nullptr, nullptr, nullptr, nullptr));
            alpha = rho / c->h_lin[0];
            cpl_update_s<<<gv, 256, 0, c->stream>>>(n, sv, r, v, alpha);
            cpl_precond<<<gb, 64, 0, c->stream>>>(q, sv, sh);
            cpl_amul<<<gc, 128, 0, c->stream>>>(q, sh, t);
            CUDA_OK(c, dots(t, sv, t, t, nullptr, nullptr, nullptr));
            omega = c->h_lin[0] / c->h_lin[1];
            cpl_update_xr<<<gv, 256, 0, c->stream>>>(n, x, r, ph, sh, sv, t, alpha, omega);
            CUDA_OK(c, dots(rw, r, nullptr, nullptr, nullptr, nullptr, r));
            rhoNext = c->h_lin[0];
            final = measure();
            c->linIterations++;
            c->launches += 7;
        } while (!stop());
    }
    CUDA_OK(c, cudaGetLastError());
    return BF_OK;
}

// Diagnostic: one Newton iteration (as bf_newton_iteration) together with the normwise backward error of its linear solve,
//     eta = ||b - A x||_2 / (||A||_F ||x||_2 + ||b||_2),
// A, b = the block system assembled from the pre-solve state by the cell-parallel assembly (long_assemble: the blocks the
// reference hands to BlockEigenSolverOF, EigenOF.C:60-175), x = the increments the sweeps produced.  Needs storeIncrements.
__global__ void linearResidualKernel(const __grid_constant__ BeamParams p, const LongParams q, double* __restrict__ acc4)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v[4] = {0, 0, 0, 0}; // r^2, b^2, A^2, x^2
    if (g < q.N) {
        const int b = q.cellBeam[g];
        const int i = (int)(g - q.cellStart[b]), n = p.nSeg[b];
        const size_t T = (size_t)(b >> 5);
        const int lane = b & 31;
        double x[3][6];
        for (int d = -1; d <= 1; d++) {
            const int ii = i + d;
            for (int k = 0; k < 6; k++) x[d + 1][k] = 0.0;
            if (ii < 0 || ii >= n) continue;
            const size_t c3 = tileRow(T, p.nmax, ii, 3, lane);
            for (int k = 0; k < 3; k++) { x[d + 1][k] = p.DW[c3 + (size_t)k * TS]; x[d + 1][3 + k] = p.DTh[c3 + (size_t)k * TS]; }
        }
        for (int r = 0; r < 6; r++) {
            double a = q.R0[6 * g + r];
            v[1] += a * a;
            for (int c = 0; c < 6; c++) {
                const double l = q.L0[36 * g + 6 * r + c], d = q.D0[36 * g + 6 * r + c], u = q.U0[36 * g + 6 * r + c];
                a -= l * x[0][c] + d * x[1][c] + u * x[2][c];
                v[2] += l * l + d * d + u * u;
            }
            v[0] += a * a;
            v[3] += x[1][r] * x[1][r];
        }
    }
    for (int k = 0; k < 4; k++) {
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if ((threadIdx.x & 31) == 0) atomicAdd(acc4 + k, v[k]);
    }
}
extern "C" int bf_newton_iteration_checked(bf_ctx* c, double sumsq[3], double* backwardError)
{
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_newton_iteration_checked between bf_evolve_begin and bf_evolve_end: the Newton loop of the open step is still running");
    if (!backwardError) return fail(c, BF_ERR_INVALID, "bf_newton_iteration_checked: null argument");
    if (!c->storeInc) return fail(c, BF_ERR_INVALID, "bf_newton_iteration_checked needs storeIncrements (the solution of the linear solve)");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (int rc = ensureLongBuffers(c)) return rc;
    {   // the assembled system of the state as it is now
        const BeamParams p = makeParams(c);
        if (p.pf) beam_point_forces<<<(unsigned)((c->B + 127) / 128), 128, 0, c->stream>>>(p);
        if (c->contrib.n) return fail(c, BF_ERR_UNSUPPORTED, "bf_newton_iteration_checked: momentum contributions are not supported");
        LongParams q;
        memset(&q, 0, sizeof q);
        q.N = c->N; q.cellBeam = c->d_cellBeam; q.cellStart = c->d_cellStart;
        q.L0 = c->d_longRows[8]; q.D0 = c->d_longRows[9]; q.U0 = c->d_longRows[10]; q.R0 = c->d_longRows[11];
        const int gb = (int)((c->N + 127) / 128);
        if (c->scheme == BF_SCHEME_NEWMARK) long_assemble<BFK_NEWMARK><<<gb, 128, 0, c->stream>>>(p, q);
        else if (c->scheme == BF_SCHEME_EULER) long_assemble<BFK_EULER><<<gb, 128, 0, c->stream>>>(p, q);
        else long_assemble<BFK_STEADY><<<gb, 128, 0, c->stream>>>(p, q);
        c->launches++;
        CUDA_OK(c, cudaGetLastError());
    }
    int rc = bf_newton_iteration(c, sumsq);
    if (rc) return rc;
    {
        const BeamParams p = makeParams(c);
        LongParams q;
        memset(&q, 0, sizeof q);
        q.N = c->N; q.cellBeam = c->d_cellBeam; q.cellStart = c->d_cellStart;
        q.L0 = c->d_longRows[8]; q.D0 = c->d_longRows[9]; q.U0 = c->d_longRows[10]; q.R0 = c->d_longRows[11];
        CUDA_OK(c, cudaMemsetAsync(c->d_sums + 4, 0, 4 * sizeof(double), c->stream));
        linearResidualKernel<<<(unsigned)((c->N + 127) / 128), 128, 0, c->stream>>>(p, q, c->d_sums + 4);
        c->launches++;
        double h[4];
        CUDA_OK(c, cudaMemcpyAsync(h, c->d_sums + 4, sizeof h, cudaMemcpyDeviceToHost, c->stream));
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
        *backwardError = sqrt(h[0]) / (sqrt(h[2]) * sqrt(h[3]) + sqrt(h[1]));
    }
    return BF_OK;
}

enum { NCCL_FLOAT64 = 8, NCCL_SUM = 0 };

// checkConvergence (Evolve.C:926-1026) on the squared global sums of iteration `iteration`.  Returns true when the
// loop ends; *status = BF_OK / BF_DIVERGED / BF_MAXITER, and the error text of the two failures.
static bool checkConvergence(bf_ctx* c, const bf_tolerances* tol, int iteration, const double s[3], double* initialResidualNorm,
                             double* currentResidualNorm, double* deltaXNorm, double* XNorm, int* status)
{
    if (!(s[0] == s[0]) || !(s[1] == s[1]) || !(s[2] == s[2])) { // NaN: a singular block -- the reference would abort inside SparseLU
        *status = BF_DIVERGED;
        fail(c, BF_DIVERGED, "Iteration %d: non-finite residual/solution norm", iteration);
        return true;
    }
    *currentResidualNorm = sqrt(s[0]);
    *deltaXNorm = sqrt(s[1]);
    *XNorm = sqrt(s[2]);
    if (iteration == 0) *initialResidualNorm = *currentResidualNorm;
    if (*currentResidualNorm <= tol->absoluteTol) return true;
    if (*currentResidualNorm <= tol->residualTol * *initialResidualNorm) return true;
    if (*deltaXNorm <= tol->solutionTol * *XNorm) return true;
    if (*currentResidualNorm >= tol->divergenceTol * *initialResidualNorm && iteration > 1) {
        fail(c, BF_DIVERGED, "Iteration %d: Diverged - Residual grew excessively.", iteration);
        *status = BF_DIVERGED;
        return true;
    }
    if (iteration >= tol->nCorrectors) {
        fail(c, BF_MAXITER, "Iteration %d: Failed - Maximum iterations reached.", iteration);
        *status = BF_MAXITER;
        return true;
    }
    return false;
}

// Elapsed times of the last finished chunk of the pipelined loop (its events have all completed: the chunk was synchronised).
static void readChunkTimes(bf_ctx* c)
{
    const int n = c->evPendingN, k0 = c->evPendingK0;
    c->evPendingN = 0;
    cudaEvent_t (*ev)[3] = c->evIter[c->evPendingSet];
    for (int j = 0; j < n; j++) {
        const int k = k0 + j;
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, ev[j][0], ev[j][1]) != cudaSuccess || cudaEventElapsedTime(&b, ev[j][1], ev[j][2]) != cudaSuccess) { cudaGetLastError(); return; }
        c->fwdMs += a; c->bwdMs += b; c->timedIters++;
        if (k < 8) {
            float g = 0;
            if (j > 0) cudaEventElapsedTime(&g, ev[j - 1][2], ev[j][0]);
            c->fwdMsK[k] += a; c->bwdMsK[k] += b; c->gapMsK[k] += g; c->itersK[k]++;
        }
    }
}

// The pipelined Newton loop: the iterations of a chunk are enqueued back to back without waiting for their norms; every
// sweep kernel looks at the global norms of the previous iteration itself (convLeave) and the ones past the end of the loop
// leave at once.  The host reads the verdicts of a whole chunk with ONE synchronisation.  The chunk is the iteration
// count of the previous step (the same on every rank, so every rank enqueues the same launches).
extern "C" int bf_evolve(bf_ctx* c, const bf_tolerances* tol, bf_step_out* out);
// One chunk is enqueued by pipeEnqueue and read back by pipeCollect; bf_evolve runs the two in a loop, bf_evolve_begin /
// bf_evolve_end put the host between them (PipeState in the context).
static void pipeInit(bf_ctx* c, const bf_tolerances* tol)
{
    PipeState& st = c->pipe;
    st.tol = *tol;
    st.initialResidualNorm = 1; st.currentResidualNorm = st.deltaXNorm = st.XNorm = BF_GREAT;
    st.status = BF_OK; st.nIter = 0; st.k0 = st.k1 = 0; st.done = false;
    c->epoch++;
    ConvCtl& cv = st.cv;
    cv.xbuf = c->d_xbuf + (size_t)(c->epoch & 1) * CONV_MAXK * c->xRanks * 4;
    cv.vflag = c->d_vflag + (size_t)(c->epoch & 1) * CONV_MAXK;
    cv.nRanks = c->xRanks; cv.iter = 0; cv.tag = 2.0 * (double)c->epoch;
    cv.absTol = tol->absoluteTol; cv.resTol = tol->residualTol; cv.solTol = tol->solutionTol; cv.divTol = tol->divergenceTol;
    cv.nCorr = tol->nCorrectors;
    cv.waitNs = c->peerWaitNs;
    st.chunk = c->lastIters < 1 ? 1 : c->lastIters > PIPE_CHUNK_MAX ? PIPE_CHUNK_MAX : c->lastIters;
}
static int pipeEnqueue(bf_ctx* c)
{
    PipeState& st = c->pipe;
    st.evSet = c->evSet;
    c->evSet ^= 1;
    st.k1 = st.k0 + st.chunk;
    if (st.k1 > st.tol.nCorrectors + 1) st.k1 = st.tol.nCorrectors + 1; // iteration nCorrectors always ends the loop
    for (int k = st.k0; k < st.k1; k++) {
        st.cv.iter = k;
        c->cvNow = st.cv;
        int rc = launchIteration(c, c->evIter[st.evSet][k - st.k0]);
        if (rc) { // leave nothing behind on the side stream
            c->cvNow.xbuf = nullptr;
            if (c->redPending) { cudaStreamWaitEvent(c->stream, c->evReduced, 0); c->redPending = false; }
            return rc;
        }
    }
    c->cvNow.xbuf = nullptr;
    if (c->redPending) { CUDA_OK(c, cudaStreamWaitEvent(c->stream, c->evReduced, 0)); c->redPending = false; }
    beam_publish_chunk<<<1, 32 * (st.k1 - st.k0), 0, c->stream>>>(st.cv, st.k0, st.k1, c->d_hres);
    c->launches++;
    CUDA_OK(c, cudaGetLastError());
    // step outputs ride behind every chunk (a chunk that turns out not to be the last is simply overwritten by the next)
    if (c->outW) { if (int rc = stagePatchPair(c, c->outEnd, c->outW, c->outTheta)) return rc; }
    readChunkTimes(c); // of the previous chunk, while the GPU works on this one
    return BF_OK;
}
static int pipeCollect(bf_ctx* c, bf_step_out* out)
{
    PipeState& st = c->pipe;
    CUDA_OK(c, cudaStreamSynchronize(c->stream));
    for (int k = st.k0; k < st.k1 && !st.done; k++) {
        const double* r = c->h_res + 4 * (k - st.k0);
        if (r[3] == 3.0) { // a peer died or never launched: its record of this iteration did not arrive (the kernels gave up waiting)
            out->status = BF_ERR_COMM; out->nIter = k;
            return fail(c, BF_ERR_COMM, "iteration %d: the norm sums of a peer rank did not arrive within %g s", k, c->peerWaitNs * 1e-9);
        }
        st.done = checkConvergence(c, &st.tol, k, r, &st.initialResidualNorm, &st.currentResidualNorm, &st.deltaXNorm, &st.XNorm, &st.status);
        st.nIter = k + 1;
        if (st.done != (r[3] == 1.0)) return fail(c, BF_ERR_CUDA, "iteration %d: host and device disagree on the convergence test", k);
    }
    if (c->timing) { c->evPendingSet = st.evSet; c->evPendingK0 = st.k0; c->evPendingN = st.nIter - st.k0; }
    st.k0 = st.k1;
    st.chunk = 1; // the prediction was short: go on one iteration at a time
    return BF_OK;
}
static int pipeFinish(bf_ctx* c, bf_step_out* out)
{
    const PipeState& st = c->pipe;
    c->lastIters = st.nIter;
    c->iOuterCorr = st.nIter;
    out->nIter = st.nIter;
    out->status = st.status;
    out->initialResidualNorm = st.initialResidualNorm;
    out->currentResidualNorm = st.currentResidualNorm;
    out->deltaXNorm = st.deltaXNorm;
    out->xNorm = st.XNorm;
    return st.status;
}
static int evolvePipelined(bf_ctx* c, const bf_tolerances* tol, bf_step_out* out)
{
    pipeInit(c, tol);
    while (!c->pipe.done) {
        if (int rc = pipeEnqueue(c)) { out->status = rc; return rc; }
        if (int rc = pipeCollect(c, out)) return rc;
    }
    return pipeFinish(c, out);
}
static bool canPipeline(const bf_ctx* c, const bf_tolerances* tol)
{
    // Pipelined loop whenever the sums can be exchanged on the device: a single context, or peers attached with
    // bf_comm_ipc_attach.  A host reducer / NCCL communicator without peer memory, the cell-parallel path and the fused
    // launch keep the host in the loop (one synchronisation per iteration).
    return !c->reducer && (!c->comm || c->xRanks > 1) && !c->longMode && !c->nPairs && (c->splitKernels || c->split || c->coopFwd || c->coopBwd) &&
           tol->nCorrectors < CONV_MAXK - 1 && !getenv("BEAMGPU_NO_PIPELINE");
}

// evolve() in two halves (include/beamgpu.h): bf_evolve_begin enqueues the Newton loop and returns, the host prepares the next
// time level (evaluates its BC series into page-locked staging, writes the previous one) while the GPU works, bf_evolve_end
// waits, checks and -- if the prediction of the iteration count was short -- runs the remaining iterations.
extern "C" int bf_evolve_begin(bf_ctx* c, const bf_tolerances* tol)
{
    if (!tol) return fail(c, BF_ERR_INVALID, "bf_evolve_begin: null argument");
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_evolve_begin: the previous bf_evolve_begin has no bf_evolve_end yet");
    CUDA_OK(c, cudaSetDevice(c->device));
    c->pipe.tol = *tol;
    c->pipe.deferred = !canPipeline(c, tol); // a host-driven loop cannot run ahead: bf_evolve_end does all of it
    if (!c->pipe.deferred) {
        c->iOuterCorr = 0;
        pipeInit(c, tol);
        if (int rc = pipeEnqueue(c)) return rc;
    }
    c->pipe.active = true;
    return BF_OK;
}
extern "C" int bf_evolve_end(bf_ctx* c, bf_step_out* out)
{
    if (!out) return fail(c, BF_ERR_INVALID, "bf_evolve_end: null argument");
    if (!c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_evolve_end without bf_evolve_begin");
    c->pipe.active = false;
    if (c->pipe.deferred) {
        const bf_tolerances tol = c->pipe.tol;
        return bf_evolve(c, &tol, out);
    }
    CUDA_OK(c, cudaSetDevice(c->device));
    memset(out, 0, sizeof *out);
    if (int rc = pipeCollect(c, out)) return rc;
    while (!c->pipe.done) {
        if (int rc = pipeEnqueue(c)) { out->status = rc; return rc; }
        if (int rc = pipeCollect(c, out)) return rc;
    }
    return pipeFinish(c, out);
}

// evolve(): Evolve.C:55-670, checkConvergence :926-1026
extern "C" int bf_evolve(bf_ctx* c, const bf_tolerances* tol, bf_step_out* out)
{
    if (!tol || !out) return fail(c, BF_ERR_INVALID, "bf_evolve: null argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    double initialResidualNorm = 1, currentResidualNorm = BF_GREAT, deltaXNorm = BF_GREAT, XNorm = BF_GREAT;
    memset(out, 0, sizeof *out);
    c->iOuterCorr = 0;
    if (c->pipe.active) return fail(c, BF_ERR_INVALID, "bf_evolve between bf_evolve_begin and bf_evolve_end");
    if (canPipeline(c, tol)) return evolvePipelined(c, tol, out);
    int status = BF_OK;
    for (;;) {
        const int iteration = c->iOuterCorr; // the value passed as iOuterCorr()++
        double s[3];
        int rc = launchIteration(c);
        if (rc) { out->status = rc; return rc; }
        if (c->comm) { // global norms: one all-reduce of three doubles per Newton iteration
            int nr = c->nccl.AllReduce(c->d_sums, c->d_sums, 3, NCCL_FLOAT64, NCCL_SUM, c->comm, c->stream);
            if (nr) { out->status = BF_ERR_COMM; return fail(c, BF_ERR_COMM, "ncclAllReduce failed: %s", c->nccl.GetErrorString(nr)); }
        }
        if ((rc = fetchSums(c, s))) { out->status = rc; return rc; }
        if (c->reducer && c->reducer(s, 3, c->reducerUser)) { out->status = BF_ERR_COMM; return fail(c, BF_ERR_COMM, "norm reducer failed"); }
        if (checkConvergence(c, tol, iteration, s, &initialResidualNorm, &currentResidualNorm, &deltaXNorm, &XNorm, &status)) break;
    }
    if (c->outW) {
        if (int rc = stagePatchPair(c, c->outEnd, c->outW, c->outTheta)) return rc;
        CUDA_OK(c, cudaStreamSynchronize(c->stream));
    }
    out->nIter = c->iOuterCorr;
    out->status = status;
    out->initialResidualNorm = initialResidualNorm;
    out->currentResidualNorm = currentResidualNorm;
    out->deltaXNorm = deltaXNorm;
    out->xNorm = XNorm;
    return status;
}

// ---- one-sided exchange over peer memory (CUDA IPC): see ConvCtl in beam_kernels.cuh -------------------------------
extern "C" int bf_comm_ipc_export(bf_ctx* c, void* handle64)
{
    if (!handle64) return fail(c, BF_ERR_INVALID, "bf_comm_ipc_export: null argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    static_assert(sizeof h == 64, "cudaIpcMemHandle_t is 64 bytes");
    CUDA_OK(c, cudaIpcGetMemHandle(&h, c->d_xbuf));
    memcpy(handle64, &h, sizeof h);
    return BF_OK;
}
extern "C" int bf_comm_ipc_attach(bf_ctx* c, int32_t nRanks, int32_t rank, const void* handles)
{
    if (!handles || nRanks < 1 || nRanks > CONV_MAX_RANKS || rank < 0 || rank >= nRanks)
        return fail(c, BF_ERR_INVALID, "bf_comm_ipc_attach: bad argument (at most %d ranks)", CONV_MAX_RANKS);
    CUDA_OK(c, cudaSetDevice(c->device));
    for (int r = 0; r < nRanks; r++) {
        if (r == rank) { c->peerXbuf[r] = c->d_xbuf; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * (size_t)r, sizeof h);
        void* ptr = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; q++) if (q != rank && c->peerXbuf[q]) { cudaIpcCloseMemHandle(c->peerXbuf[q]); c->peerXbuf[q] = nullptr; }
            c->peerXbuf[rank] = c->d_xbuf;
            return fail(c, BF_ERR_COMM, "cudaIpcOpenMemHandle of rank %d failed: %s", r, cudaGetErrorString(e));
        }
        c->peerXbuf[r] = (double*)ptr;
    }
    c->xRanks = nRanks;
    c->xRank = rank;
    c->peerXbuf[rank] = c->d_xbuf;
    return BF_OK;
}

extern "C" int bf_set_reducer(bf_ctx* c, bf_reduce_fn fn, void* user)
{
    c->reducer = fn;
    c->reducerUser = user;
    return BF_OK;
}

// ---- NCCL (dlopen()ed so that single-GPU use has no NCCL dependency) ----------
static int loadNccl(NcclApi* api, char* err)
{
    if (api->handle) return BF_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        api->handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (api->handle) break;
    }
    if (!api->handle) { snprintf(err, 512, "cannot dlopen libnccl.so.2: %s", dlerror()); return BF_ERR_COMM; }
    api->GetUniqueId = (int (*)(void*))dlsym(api->handle, "ncclGetUniqueId");
    api->CommInitRank = (int (*)(void**, int, Id128, int))dlsym(api->handle, "ncclCommInitRank");
    api->AllReduce = (int (*)(const void*, void*, size_t, int, int, void*, cudaStream_t))dlsym(api->handle, "ncclAllReduce");
    api->CommDestroy = (int (*)(void*))dlsym(api->handle, "ncclCommDestroy");
    api->GetErrorString = (const char* (*)(int))dlsym(api->handle, "ncclGetErrorString");
    if (!api->GetUniqueId || !api->CommInitRank || !api->AllReduce || !api->CommDestroy || !api->GetErrorString) {
        snprintf(err, 512, "libnccl is missing a required symbol");
        return BF_ERR_COMM;
    }
    return BF_OK;
}

extern "C" int bf_nccl_unique_id(void* id128)
{
    static NcclApi api;
    if (!id128) return BF_ERR_INVALID;
    if (loadNccl(&api, g_create_err)) return BF_ERR_COMM;
    return api.GetUniqueId(id128) == 0 ? BF_OK : BF_ERR_COMM;
}

extern "C" int bf_comm_init_nccl(bf_ctx* c, const void* id128, int32_t nRanks, int32_t rank)
{
    if (!id128 || nRanks < 1 || rank < 0 || rank >= nRanks) return fail(c, BF_ERR_INVALID, "bf_comm_init_nccl: bad argument");
    CUDA_OK(c, cudaSetDevice(c->device));
    if (loadNccl(&c->nccl, c->err)) return BF_ERR_COMM;
    Id128 id;
    memcpy(&id, id128, sizeof id);
    int nr = c->nccl.CommInitRank(&c->comm, nRanks, id, rank);
    if (nr) { c->comm = nullptr; return fail(c, BF_ERR_COMM, "ncclCommInitRank failed: %s", c->nccl.GetErrorString(nr)); }
    return BF_OK;
}
