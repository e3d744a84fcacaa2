/* beamgpu.h — C-ABI boundary of the B200-native beamFoam Newton-iteration path.
 *
 * This header is the drop-in boundary between an OpenFOAM host (the beamModel
 * subclass `coupledTotalLagNewtonRaphsonBeamGPU`) and the
 * sm_100a CUDA implementation of ONE hot path of solids4foam/beamFoam:
 * the per-Newton-iteration update of `coupledTotalLagNewtonRaphsonBeam`.
 *
 * Reference interfaces replaced (paths relative to the reference checkout,
 * CTL = src/wireBunchingModels/beamModels/coupledTotalLagNewtonRaphsonBeam):
 *   scalar evolve()                       CTL/coupledTotalLagNewtonRaphsonBeamEvolve.C:55-670
 *   void   updateEqnCoefficients()        CTL/...Evolve.C:673-753
 *   void   assembleMatrixCoefficients()   CTL/...Matrix.C:55-475
 *   void   assembleBoundaryConditions()   CTL/...BoundaryConditions.C:64-1190
 *   BlockEigenSolverOF::solve()           src/wireBunchingModels/numerics/BlockEigenSolvers/
 *                                         BlockEigenSolverOF/BlockEigenSolverOF.C:200-412
 *   void   updateSolutionVariables()      CTL/...Evolve.C:757-923
 *   bool   checkConvergence()             CTL/...Evolve.C:926-1026
 *   fvPatchField evaluate()/updateCoeffs  src/wireBunchingModels/beamModels/fvPatchFields/<name>
 *
 * The same header is implemented twice:
 *   beamfoam_b200/libbeamgpu.so   — the product: hand-written FP64 CUDA, sm_100a
 *   oracle/_build/libbeamoracle.so — TEST INFRASTRUCTURE ONLY: a literal CPU
 *                                    restatement of the reference arithmetic
 *
 * Conventions
 *   - POD only: plain pointers and sizes; no C++ / torch types cross the ABI.
 *   - All floating point is IEEE FP64; indices are 32-bit (OpenFOAM `label`)
 *     except global counts, which are int64_t.
 *   - Every function returns an int status (BF_OK == 0).  Nothing aborts or
 *     throws across the boundary; bf_last_error() gives the message.  The
 *     reference's FatalError cases (divergence, max iterations: Evolve.C:1003-1022)
 *     map to BF_DIVERGED / BF_MAXITER.
 *   - Host arrays use the reference's own field representation and the
 *     createBeamMesh ordering (applications/utilities/createBeamMesh/createBeamMesh.C:199-408):
 *       cells    : beam-major contiguous, cellStart[k] = sum(nSeg[<k])
 *       internal faces : beam-major contiguous, nSeg[k]-1 per beam,
 *                        owner = i-1+cellStart[k], neighbour = i+cellStart[k]
 *       boundary : patch left_k (1 face, faceCell = first cell of beam k),
 *                  patch right_k (1 face, faceCell = last cell of beam k)
 *       vector   : 3 contiguous doubles (x y z)
 *       tensor   : 9 contiguous doubles (xx xy xz yx yy yz zx zy zz)
 *   - One context per host thread; one CUDA device per context.  Calls are
 *     synchronous at return.
 */
#ifndef BEAMGPU_H
#define BEAMGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes -------------------------------------------------------- */
#define BF_OK               0
#define BF_ERR_INVALID      1   /* bad argument / inconsistent sizes             */
#define BF_ERR_CUDA         2   /* CUDA runtime error (message in bf_last_error) */
#define BF_ERR_NOMEM        3
#define BF_DIVERGED         4   /* Evolve.C:1003-1012 "Diverged - Residual grew excessively" */
#define BF_MAXITER          5   /* Evolve.C:1015-1022 "Failed - Maximum iterations reached"  */
#define BF_ERR_UNSUPPORTED  6   /* e.g. coupled patches: BoundaryConditions.C:79-83 notImplemented */
#define BF_ERR_COMM         7   /* NCCL / reducer failure */

/* ---- enumerations -------------------------------------------------------- */
/* d2dt2Schemes default (CTL/coupledTotalLagNewtonRaphsonBeam.C:668-683,708-765) */
#define BF_SCHEME_STEADY    0
#define BF_SCHEME_NEWMARK   1
#define BF_SCHEME_EULER     2

/* W boundary-condition classes (all derive from fixedValueFvPatchVectorField) */
#define BF_W_FIXED          0   /* fixedValue, fixedDisplacement                      */
#define BF_W_FORCE          1   /* forceBeamDisplacementNR                            */
#define BF_W_FOLLOWER       2   /* followerForceBeamDisplacementNR (material-frame F) */
#define BF_W_SPRING         3   /* linearSpringForceBeamDisplacementNR                */

/* Theta boundary-condition classes */
#define BF_TH_FIXED         0   /* fixedValue, fixedRotation */
#define BF_TH_MOMENT        1   /* momentBeamRotationNR      */

/* patch ends */
#define BF_END_LEFT         0
#define BF_END_RIGHT        1

/* per-step boundary values set by bf_set_bc_values (what the reference's
 * updateCoeffs() evaluates from its time series at the NEW time) */
#define BF_BCV_W            0   /* prescribed W_b      (fixedValue / displacementSeries)   */
#define BF_BCV_THETA        1   /* prescribed Theta_b  (fixedValue / thetaSeries)          */
#define BF_BCV_FORCE        2   /* force(): forceSeries; followerForceSeries (material
                                   frame); offsetForceSeries for the spring BC             */
#define BF_BCV_MOMENT       3   /* moment(): momentSeries                                  */

/* field identifiers for bf_set_field / bf_get_field */
#define BF_FIELD_W          0   /* volVectorField  W        + boundary */
#define BF_FIELD_THETA      1   /* volVectorField  Theta    + boundary */
#define BF_FIELD_LAMBDA     2   /* volTensorField  Lambda   + boundary */
#define BF_FIELD_DW         3   /* volVectorField  DW       + boundary (last increment) */
#define BF_FIELD_DTHETA     4   /* volVectorField  DTheta   + boundary */
#define BF_FIELD_U          5   /* volVectorField  U        (internal only) */
#define BF_FIELD_ACCL       6   /* volVectorField  Accl     (internal only) */
#define BF_FIELD_OMEGA      7   /* volVectorField  Omega    (internal only) */
#define BF_FIELD_DOTOMEGA   8   /* volVectorField  dotOmega (internal only) */
#define BF_FIELD_LAMBDAF    9   /* surfaceTensorField Lambdaf */
#define BF_FIELD_GAMMA      10  /* surfaceVectorField Gamma   */
#define BF_FIELD_K          11  /* surfaceVectorField K       */
#define BF_FIELD_Q          12  /* surfaceVectorField Q       */
#define BF_FIELD_M          13  /* surfaceVectorField M       */
#define BF_FIELD_COUNT      14

typedef struct bf_ctx bf_ctx;

/* ---- construction -------------------------------------------------------- */

/* Mesh of createBeamMesh: nBeams straight prism chains along global x
 * (createBeamMesh.C:100-135,199-408).  dZ = length/nSeg; internal
 * deltaCoeffs = 1/dZ, boundary deltaCoeffs = 2/dZ, weights = 0.5. */
typedef struct bf_layout {
    int64_t        nBeams;
    const int32_t* nSeg;        /* [nBeams] cells per beam ("nSegments")            */
    const double*  length;      /* [nBeams] beam length                              */
    const double*  cellLength;  /* [sum(nSeg)] L() per cell, or NULL => dZ of its beam
                                   (CTL.C:1060-1108: HermiteSpline::segLengths())    */
} bf_layout;

/* Per-beam constants the reference builds in its constructor
 * (CTL.C:538-667,781-845; beamModel.H:449-584). */
typedef struct bf_beam_props {
    const double* CQ;          /* [nBeams*3] diag(EA, GA, GA)                        */
    const double* CM;          /* [nBeams*3] diag(GJ, E*Iyy, E*Izz)                  */
    const double* ARho;        /* [nBeams]   rho*A            (NULL => 0)            */
    const double* CIRho;       /* [nBeams*3] kCI*rho*(J, Iyy, Izz) (NULL => 0)       */
    const double* weight;      /* [nBeams*3] rho*A*g  (Evolve.C:201-209) (NULL => 0) */
    const double* refLambdaf;  /* [nBeams*9] rigid initial rotation T written by
                                  setInitialPositionBeam (refLambdaf = T, dR0Ds = T.x,
                                  sign-flipped on the left patch: CTL.C:987-994);
                                  NULL => identity                                   */
} bf_beam_props;

/* Boundary-condition classes per beam end; index = end*nBeams + beam. */
typedef struct bf_bc {
    const int32_t* wType;            /* [2*nBeams] BF_W_*                            */
    const int32_t* thetaType;        /* [2*nBeams] BF_TH_*                           */
    const double*  springStiffness;  /* [2*nBeams]   (NULL if no BF_W_SPRING)        */
    const double*  springDirection;  /* [2*nBeams*3] normalised by the library as in
                                        linearSpringForce...C:116-119                */
} bf_bc;

typedef struct bf_options {
    int32_t scheme;          /* BF_SCHEME_*                                          */
    double  newmarkBeta;     /* default 0.25  (CTL.C:755)                            */
    double  newmarkGamma;    /* default 0.5   (CTL.C:756)                            */
    int32_t device;          /* CUDA device ordinal (ignored by the oracle)          */
    int32_t storeIncrements; /* !=0: keep the DW/DTheta cell fields in memory so that
                                bf_get_field(BF_FIELD_DW/DTHETA) returns them        */
    int32_t reserved[6];     /* must be zero */
} bf_options;

/* Keys of <model>Coeffs in constant/beamProperties (Evolve.C:61-102). */
typedef struct bf_tolerances {
    int32_t nCorrectors;     /* default 1000 */
    double  residualTol;     /* rtol; default 1e-8 (falls back to convergenceTol)    */
    double  solutionTol;     /* stol; default 1e-10 */
    double  absoluteTol;     /* atol; default 1e-30 */
    double  divergenceTol;   /* divtol; default 1e6 */
} bf_tolerances;

/* Result of one evolve(): iOuterCorr() after the loop and the norms of the
 * last checkConvergence call (Evolve.C:104-107,644-669). */
typedef struct bf_step_out {
    int32_t nIter;               /* number of linear solves performed this step      */
    int32_t status;              /* BF_OK, BF_DIVERGED or BF_MAXITER                 */
    double  initialResidualNorm; /* what evolve() returns                            */
    double  currentResidualNorm;
    double  deltaXNorm;
    double  xNorm;
} bf_step_out;

/* Sum-reduction of `n` doubles over all ranks that hold a shard of the bundle
 * (in place).  Return 0 on success.  An OpenFOAM host passes a wrapper of
 * Pstream `reduce(..., sumOp<scalar>())`; a torch host passes an all_reduce. */
typedef int (*bf_reduce_fn)(double* values, int32_t n, void* user);

/* Replaces: coupledTotalLagNewtonRaphsonBeam constructor state set-up
 * (CTL.C:74-1182) for the quantities the hot path touches.  Fresh-run initial
 * values: W=Theta=0, Lambda=Lambdaf=I, Gamma=K=Q=M=0,
 * U=Accl=Omega=dotOmega=0, prevIter = current. */
int bf_create(const bf_layout* layout, const bf_beam_props* props,
              const bf_bc* bc, const bf_options* options, bf_ctx** out);
int bf_destroy(bf_ctx* ctx);

const char* bf_last_error(const bf_ctx* ctx);   /* ctx may be NULL: last create error */
const char* bf_version(void);
/* "cuda-sm100a" for libbeamgpu, "cpu-oracle" for the oracle */
const char* bf_backend(void);

int64_t bf_num_cells(const bf_ctx* ctx);
int64_t bf_num_internal_faces(const bf_ctx* ctx);
int64_t bf_num_beams(const bf_ctx* ctx);

/* ---- state mirroring (host <-> authoritative solver state) ------------------
 * Replaces the field members of CTL.H:108-177 being read/written by the host.
 * Vol fields:     internal [nCells*nc], left/right = patch values [nBeams*nc].
 * Surface fields: internal [nInternalFaces*nc], left/right [nBeams*nc].
 * Any pointer may be NULL (that part is skipped).  Setting W or Theta also
 * refreshes their prevIter copy (storePrevIter, CTL.C:1110-1111).
 *
 * The host always sees the reference's representation; libbeamgpu keeps some fields
 * in a compact form and converts at this boundary (libbeamoracle stores them as is):
 *  - Lambda (internal) and Lambdaf are rotation tensors and are kept as unit
 *    quaternions (Lambdaf as the product Lambdaf & refLambdaf): what is set must be a
 *    proper rotation, what is read back equals it to rounding (1e-15), not bit for bit.
 *  - Gamma is a pure function of W and Lambdaf (Evolve.C:881-887) and is evaluated on
 *    request; setting it is accepted and ignored.
 *  - U and Omega are kept as the Newmark step constants U - gamma*deltaT*Accl,
 *    Omega - gamma*deltaT*dotOmega; setting Accl / dotOmega leaves U / Omega unchanged,
 *    as in the reference. */
int bf_set_field(bf_ctx* ctx, int32_t field, const double* internal,
                 const double* left, const double* right);
int bf_get_field(bf_ctx* ctx, int32_t field, double* internal,
                 double* left, double* right);

/* Asynchronous mirror of the INTERNAL parts of nFields fields (reference ordering) into host
 * buffers, which should be pinned (bf_host_alloc): returns once the copies are queued; they
 * overlap whatever the context does next (the Newton loop of the following step).
 * bf_mirror_wait blocks until the host buffers are complete.  At most 9 doubles per cell in
 * total per call (e.g. W + Theta + one more vector field).  Synchronous transfers of INTERNAL
 * fields (bf_set_field / bf_get_field) drain a pending mirror first; patch-value transfers
 * (bf_set_bc_values, boundary-only bf_get_field) do not. */
int bf_mirror_begin(bf_ctx* ctx, int32_t nFields, const int32_t* fields, double* const* hostInternal);
int bf_mirror_wait(bf_ctx* ctx);

/* Distributed loads q, m (volVectorFields "q","m"; Evolve.C:193-198,274-279).
 * [nCells*3] each; NULL => zero. */
int bf_set_distributed_loads(bf_ctx* ctx, const double* q, const double* m);

/* constant/pointForces (beamModel.C:863-866; Evolve.C:211-271): n point forces, each on cell
 * cell[k] (index into the cells of this context) at relative coordinate zeta[k] in [-1, 1]
 * with force[3k..] = the value of its time series at the current time.  Replaces the list
 * set by a previous call; n = 0 removes all point forces. */
int bf_set_point_forces(bf_ctx* ctx, int64_t n, const int64_t* cell, const double* zeta,
                        const double* force);

/* constant/beamMomentumContributionProperties: the run-time selectable explicit momentum sources
 * (src/wireBunchingModels/beamMomentumContribution; created in CTL.C:1137-1182, added to the source in
 * Evolve.C:539-571 for every d2dt2Scheme; their diagCoeff() is zero).  Both evaluate, per cell, the
 * mid-point derivative of the Hermite spline through the current face points and tangents
 * (numerics/HermiteSpline/HermiteSpline.C:83-108,300-385,468-513; CTL.C:1263-1427,1518-1645) and the
 * cell velocity U:
 *   BF_CONTRIB_MORISON_DRAG    morisonDragContribution.C:103-195   coeffs = { Cdn, Cdt, rhoFluid }
 *   BF_CONTRIB_GROUND_CONTACT  groundContactContribution.C:111-236 coeffs = { kNormal, cNormal,
 *                              kTangential, muFriction, groundZ }
 * radius[nBeams] = crossSection.R() (circle.H:117, rectangle.H:133: h/2); translation[nBeams*3] = the
 * -translate vector of setInitialPositionBeam (refW = T & C + translate - C; NULL => 0).
 * The reference evaluates the spline of beam 0 only (currentBeamPoints() with the default bI = 0), so its
 * result is defined for single-beam cases; here every beam gets its own spline.  morisonDrag with the
 * steadyState scheme is a FatalError in the reference: BF_ERR_UNSUPPORTED.  n = 0 removes all. */
#define BF_CONTRIB_MORISON_DRAG    0
#define BF_CONTRIB_GROUND_CONTACT  1
#define BF_MAX_CONTRIBUTIONS       8
typedef struct bf_momentum_contribution {
    int32_t type;
    int32_t reserved;
    double  coeffs[6];
} bf_momentum_contribution;
int bf_set_momentum_contributions(bf_ctx* ctx, int32_t n, const bf_momentum_contribution* list,
                                  const double* radius, const double* translation);

/* Per-step boundary values == what W_/Theta_.boundaryFieldRef().updateCoeffs()
 * (Evolve.C:156-157) evaluates from the BC time series at the new time.
 * values: [nBeams*3] for patch `end`; entries of beams whose BC class does not
 * use `kind` are ignored. */
int bf_set_bc_values(bf_ctx* ctx, int32_t end, int32_t kind, const double* values);
/* The same without the wait for the upload: `values` must be page-locked (bf_host_alloc)
 * and must stay unchanged until the next bf_evolve / bf_newton_iteration has returned
 * (a host that refills the same buffer once per time step satisfies this by construction). */
int bf_set_bc_values_async(bf_ctx* ctx, int32_t end, int32_t kind, const double* values);
/* Patch values of W and Theta at one end of every beam ([nBeams][3] each) in one
 * call -- the per-step read of functionObjects/beamDisplacements
 * (functionObjects/beamDisplacements/beamDisplacements.C) -- equal to
 * bf_get_field(W, ...) + bf_get_field(THETA, ...) for that side. */
int bf_get_patch_pair(bf_ctx* ctx, int32_t end, double* W, double* Theta);
/* Step outputs: registers two page-locked host arrays ([nBeams][3] each, bf_host_alloc) that
 * every following bf_evolve fills with the patch values of W and Theta of `end` BEFORE it
 * returns -- the same values bf_get_patch_pair would deliver, riding behind the Newton loop
 * on the context's stream, so that a host that reads them after every step
 * (functionObjects/beamDisplacements) pays no second synchronisation.  W = Theta = NULL
 * unregisters. */
int bf_set_step_outputs(bf_ctx* ctx, int32_t end, double* W, double* Theta);

/* Coupled bundles: inter-beam couplings turn the union of block-tridiagonal systems into ONE block-sparse
 * system -- the shape of fvBlockMatrices/multibeamFvBlockMatrix/multibeamFvBlockMatrix.C:298-470 (per-beam tridiagonal blocks
 * plus 6x6 blocks between cells of different beams) -- which is solved by the preconditioned BiCGStab of
 * numerics/blockLduMatrix/BlockLduSolvers/BlockBiCGStab/myBlockBiCGStabSolver.C:79-189 with the per-beam block-Thomas solve
 * as the preconditioner.  The reference compiles neither and its contact model is not in its repository, so the coupling is
 * SYNTHETIC and parity is pinned to the CPU oracle only: pair p is a linear spring of stiffness[p] between the centres of cells
 * cellA[p] and cellB[p] (indices of this context, createBeamMesh order), force on A = k (W_B - W_A).  nPairs = 0 restores
 * the uncoupled path.  bf_set_linear_solver: maxIter / tolerance / relTol of the fvSolution solver dictionary
 * (BlockLduSolver::stop of foam-extend 4.1: cmptMax of the componentwise L1 residual over gSum(cmptMag(b))); defaults
 * 1000, 1e-13, 0.  bf_linear_iterations: BiCGStab iterations of the last Newton iteration. */
int bf_set_couplings(bf_ctx* ctx, int64_t nPairs, const int64_t* cellA, const int64_t* cellB, const double* stiffness);
int bf_set_linear_solver(bf_ctx* ctx, int32_t maxIter, double tolerance, double relTol);
int32_t bf_linear_iterations(const bf_ctx* ctx);

/* ---- the hot path ---------------------------------------------------------- */

/* The time index advanced (runTime++ of beamFoam.C:73-79): sets deltaT, resets
 * iOuterCorr and opens a new old-time level.  The Newmark predictor of
 * Evolve.C:167-186 (Euler: the capture of W/U/Omega/Lambda.oldTime()) is applied
 * by the FIRST Newton iteration after this call and by no other: bf_evolve /
 * bf_newton_iteration called again without a new bf_begin_step (outer FSI
 * iterations, a retry with other tolerances) continue from the current state of
 * the same time level.  Call it once per time step, before bf_evolve.  A step
 * that ended BF_DIVERGED leaves the state unusable (no roll-back): reset the
 * fields with bf_set_field before going on. */
int bf_begin_step(bf_ctx* ctx, double deltaT);

/* One pass of the do{}-body of evolve() (Evolve.C:119-632): coefficients,
 * assembly, BC closure, loads+inertia, block-tridiagonal solve, state update,
 * BC evaluate.  sumsq[0] = sum(source^2) (= currentResidualNorm^2, x0 = 0),
 * sumsq[1] = sum(solVec^2), sumsq[2] = sum|W|^2 + sum|Theta|^2, LOCAL to this
 * context (EigenOF.C:279-281 "use sum instead of gSum"); no reducer applied. */
int bf_newton_iteration(bf_ctx* ctx, double sumsq[3]);

/* Diagnostic twin of bf_newton_iteration: the same iteration, plus the normwise backward error of its linear solve
 *   eta = ||b - A x||_2 / (||A||_F ||x||_2 + ||b||_2)
 * over the whole bundle, with A, b the block system of EigenOF.C:60-175 assembled from the pre-solve state by an independent
 * kernel and x the increments the solve produced (the context must keep them: bf_options.storeIncrements).  A backward
 * stable solve gives eta of the order of 1e-16; tests/test_gpu_robustness.py reports it for every kernel family. */
int bf_newton_iteration_checked(bf_ctx* ctx, double sumsq[3], double* backwardError);

/* The whole Newton loop of evolve() with checkConvergence (Evolve.C:926-1026)
 * evaluated on the (reduced) norms.  Always fills `out`. */
int bf_evolve(bf_ctx* ctx, const bf_tolerances* tol, bf_step_out* out);

/* bf_evolve in two halves.  beamFoam.C:73-89 runs one time step after the other: updateCoeffs() of the boundary conditions
 * (Evolve.C:156-157), the Newton loop, the function objects, runTime.write().  The BC series of the NEXT time level depend on
 * nothing the solve produces, so a host can evaluate them (into page-locked staging) and do its writing while the GPU is
 * busy: bf_evolve_begin enqueues the Newton loop of the open time step and returns at once, bf_evolve_end waits for it,
 * evaluates checkConvergence on its norms exactly as bf_evolve does (running further iterations if the loop was not over)
 * and fills `out`; the pair is bf_evolve, result for result.  Between the two calls the context accepts no call that would
 * queue work behind the loop (bf_begin_step, bf_set_bc_values*, bf_set_field / bf_get_field, bf_mirror_begin, bf_get_patch_pair,
 * the load setters, bf_newton_iteration, bf_evolve ... return BF_ERR_INVALID): hand
 * the prepared values over after bf_evolve_end.  Where the loop needs the host in every iteration (a reducer callback,
 * NCCL without peer memory, the cell-parallel path) bf_evolve_begin only records the request and bf_evolve_end does the work. */
int bf_evolve_begin(bf_ctx* ctx, const bf_tolerances* tol);
int bf_evolve_end(bf_ctx* ctx, bf_step_out* out);

/* Multi-GPU: beams are sharded across contexts; the only exchange is the sum of
 * the three squared norms per Newton iteration. */
int bf_set_reducer(bf_ctx* ctx, bf_reduce_fn fn, void* user);
/* Built-in reducer: ncclAllReduce(sum, 3 x double) on the context's stream.
 * uniqueId = the 128 bytes of ncclUniqueId created on rank 0 and broadcast by the
 * host (MPI_Bcast in an OpenFOAM host).  libnccl is dlopen()ed on first use. */
int bf_nccl_unique_id(void* uniqueId128);
int bf_comm_init_nccl(bf_ctx* ctx, const void* uniqueId128, int32_t nRanks, int32_t rank);
/* One-sided exchange over peer memory (NVLink / NVSwitch, one process per GPU on one node): every context owns a small
 * exchange buffer; bf_comm_ipc_export returns its 64-byte CUDA IPC handle, the host gathers the handles of all ranks
 * (MPI_Allgather in an OpenFOAM host) and hands the nRanks x 64 bytes, in rank order, to bf_comm_ipc_attach.  From then on
 * the reduction kernel of a Newton iteration writes this rank's three sums straight into every rank's buffer, every sweep
 * kernel evaluates checkConvergence (Evolve.C:926-1026) on the device, and bf_evolve enqueues whole Newton loops without a
 * host round trip or a collective per iteration (the same device-side loop serves a single context).  Results are
 * identical to the NCCL / reducer path: the records are added in rank order on every rank.  BF_ERR_COMM if the peers
 * cannot be mapped (the context then keeps its previous reducer). */
int bf_comm_ipc_export(bf_ctx* ctx, void* handle64);
int bf_comm_ipc_attach(bf_ctx* ctx, int32_t nRanks, int32_t rank, const void* handles);

/* ---- observables computed next to the state ----------------------------------
 * Replaces functionObjects/beamEnergyData/beamEnergyData.C:84-215 (a reduction over every cell and face
 * per write step; here it does not need the fields on the host).  out[3*b + 0] internal (strain) energy,
 * + 1 linear kinetic energy, + 2 angular kinetic energy of beam b of this context. */
int bf_beam_energy(bf_ctx* ctx, double* out);

/* ---- instrumentation --------------------------------------------------------
 * Kernel launches issued by this context since creation (bench gpu_launches),
 * and device time in ms of the Newton-iteration kernels since the last reset,
 * measured with CUDA events on the context's stream.  The product path is two launches per
 * iteration (beam_forward, beam_backward): fwd_ms and bwd_ms hold their times.  With
 * BEAMGPU_FUSED_KERNEL=1 in the environment both sweeps run in one launch (beam_newton); fwd_ms
 * then holds its time and bwd_ms is 0. */
int64_t bf_launch_count(const bf_ctx* ctx);
/* Which kernel family bf_create selected for this bundle (diagnostic; the results do
 * not depend on it beyond rounding): BF_FAMILY_PER_BEAM one thread per beam (bundles of
 * at least about one wave of beams), BF_FAMILY_GROUP assembler + eliminator warps (small bundles),
 * BF_FAMILY_SPLIT two threads per beam that eliminate from both ends and meet in the
 * middle, BF_FAMILY_CELL one thread per cell with block cyclic reduction (few long beams),
 * BF_FAMILY_CPU for an implementation without kernels (the test oracle).  Mixed pairs
 * forced through BEAMGPU_COOP=fwd|bwd report the family of the forward sweep. */
#define BF_FAMILY_CPU (-1)
#define BF_FAMILY_PER_BEAM 0
#define BF_FAMILY_GROUP 1
#define BF_FAMILY_CELL 2
#define BF_FAMILY_SPLIT 3
#define BF_FAMILY_SPLIT_CELL 4 /* split forward sweep, backward sweep with one thread per cell (small uniform bundles) */
#define BF_FAMILY_SPLIT_GROUP 5 /* assembler + eliminator warps per HALF beam, backward sweep with one thread per cell (bundles below half a wave) */
int32_t bf_kernel_family(const bf_ctx* ctx);
/* Split families: the number of cells of the left part of every beam (the two threads of a beam meet at the face after it);
 * 0 otherwise.  BEAMGPU_SPLIT_AT in the environment of bf_create forces it (diagnostic: results depend on it at rounding level). */
int32_t bf_split_point(const bf_ctx* ctx);
int bf_kernel_time_ms(bf_ctx* ctx, int32_t reset, double* fwd_ms, double* bwd_ms,
                      int64_t* iterations);
int bf_set_timing(bf_ctx* ctx, int32_t enabled);
/* Device-side stopwatch for a caller-chosen region: a pair of CUDA events recorded on the
 * context's stream (which = 0 start, 1 stop); elapsed waits for the stop event. */
int bf_region_mark(bf_ctx* ctx, int32_t which);
int bf_region_elapsed_ms(bf_ctx* ctx, double* ms);

/* Pinned host memory for the mirror buffers (cudaHostAlloc; malloc in the oracle). */
int bf_host_alloc(void** ptr, int64_t bytes);
int bf_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* BEAMGPU_H */
