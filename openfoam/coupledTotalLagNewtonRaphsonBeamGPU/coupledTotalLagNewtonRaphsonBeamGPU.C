/*---------------------------------------------------------------------------*\
    coupledTotalLagNewtonRaphsonBeamGPU — OpenFOAM-side binding of include/beamgpu.h.
    Shipped as source (no OpenFOAM in this repository's build environment).
\*---------------------------------------------------------------------------*/

#include "coupledTotalLagNewtonRaphsonBeamGPU.H"
#include "addToRunTimeSelectionTable.H"
#include "forceBeamDisplacementNRFvPatchVectorField.H"
#include "followerForceBeamDisplacementNRFvPatchVectorField.H"
#include "linearSpringForceBeamDisplacementNRFvPatchVectorField.H"
#include "momentBeamRotationNRFvPatchVectorField.H"
#include "Pstream.H"

namespace Foam
{
namespace beamModels
{

defineTypeNameAndDebug(coupledTotalLagNewtonRaphsonBeamGPU, 0);
addToRunTimeSelectionTable(beamModel, coupledTotalLagNewtonRaphsonBeamGPU, dictionary);

// * * * * * * * * * * * * * Private Member Functions  * * * * * * * * * * * //

void coupledTotalLagNewtonRaphsonBeamGPU::check(const int status, const char* where) const
{
    if (status != BF_OK)
    {
        FatalErrorInFunction
            << where << ": " << bf_last_error(ctx_) << " (status " << status << ")"
            << abort(FatalError);
    }
}

int coupledTotalLagNewtonRaphsonBeamGPU::reduceNorms(double* v, int32_t n, void*)
{
    for (int32_t i = 0; i < n; i++) { reduce(v[i], sumOp<scalar>()); }
    return 0;
}

// Fields are contiguous arrays of vector (3 doubles) / tensor (9 doubles, xx xy xz ...), cells and
// internal faces in createBeamMesh order: exactly the host layout of beamgpu.h.  Patch k (left_k /
// right_k, or left/right) holds one face per beam.
#define PATCH_PTRS(f, n, L, R)                                                              \
    List<scalar> L(n*nBeams()), R(n*nBeams());                                              \
    for (label b = 0; b < nBeams(); b++)                                                    \
    {                                                                                       \
        const auto& pl = f.boundaryField()[startPatchIndex(b)];                             \
        const auto& pr = f.boundaryField()[endPatchIndex(b)];                               \
        const label il = nBeams() > 1 ? 0 : b;                                              \
        for (label k = 0; k < n; k++) { L[n*b + k] = pl[il][k]; R[n*b + k] = pr[il][k]; }   \
    }

void coupledTotalLagNewtonRaphsonBeamGPU::upload(const label id, const volVectorField& f)
{
    PATCH_PTRS(f, 3, L, R)
    check(bf_set_field(ctx_, id, reinterpret_cast<const double*>(f.primitiveField().cdata()), L.cdata(), R.cdata()), "bf_set_field");
}
void coupledTotalLagNewtonRaphsonBeamGPU::upload(const label id, const volTensorField& f)
{
    PATCH_PTRS(f, 9, L, R)
    check(bf_set_field(ctx_, id, reinterpret_cast<const double*>(f.primitiveField().cdata()), L.cdata(), R.cdata()), "bf_set_field");
}
void coupledTotalLagNewtonRaphsonBeamGPU::upload(const label id, const surfaceVectorField& f)
{
    PATCH_PTRS(f, 3, L, R)
    check(bf_set_field(ctx_, id, reinterpret_cast<const double*>(f.primitiveField().cdata()), L.cdata(), R.cdata()), "bf_set_field");
}
void coupledTotalLagNewtonRaphsonBeamGPU::upload(const label id, const surfaceTensorField& f)
{
    PATCH_PTRS(f, 9, L, R)
    check(bf_set_field(ctx_, id, reinterpret_cast<const double*>(f.primitiveField().cdata()), L.cdata(), R.cdata()), "bf_set_field");
}

#define UNPACK_PATCHES(f, n, L, R)                                                          \
    for (label b = 0; b < nBeams(); b++)                                                    \
    {                                                                                       \
        auto& pl = f.boundaryFieldRef()[startPatchIndex(b)];                                \
        auto& pr = f.boundaryFieldRef()[endPatchIndex(b)];                                  \
        const label il = nBeams() > 1 ? 0 : b;                                              \
        for (label k = 0; k < n; k++) { pl[il][k] = L[n*b + k]; pr[il][k] = R[n*b + k]; }   \
    }

void coupledTotalLagNewtonRaphsonBeamGPU::download(const label id, volVectorField& f, const bool boundary)
{
    List<scalar> L(3*nBeams()), R(3*nBeams());
    check(bf_get_field(ctx_, id, reinterpret_cast<double*>(f.primitiveFieldRef().data()),
                       boundary ? L.data() : nullptr, boundary ? R.data() : nullptr), "bf_get_field");
    if (boundary) { UNPACK_PATCHES(f, 3, L, R) }
}
void coupledTotalLagNewtonRaphsonBeamGPU::download(const label id, volTensorField& f)
{
    List<scalar> L(9*nBeams()), R(9*nBeams());
    check(bf_get_field(ctx_, id, reinterpret_cast<double*>(f.primitiveFieldRef().data()), L.data(), R.data()), "bf_get_field");
    UNPACK_PATCHES(f, 9, L, R)
}
void coupledTotalLagNewtonRaphsonBeamGPU::download(const label id, surfaceVectorField& f)
{
    List<scalar> L(3*nBeams()), R(3*nBeams());
    check(bf_get_field(ctx_, id, reinterpret_cast<double*>(f.primitiveFieldRef().data()), L.data(), R.data()), "bf_get_field");
    UNPACK_PATCHES(f, 3, L, R)
}
void coupledTotalLagNewtonRaphsonBeamGPU::download(const label id, surfaceTensorField& f)
{
    List<scalar> L(9*nBeams()), R(9*nBeams());
    check(bf_get_field(ctx_, id, reinterpret_cast<double*>(f.primitiveFieldRef().data()), L.data(), R.data()), "bf_get_field");
    UNPACK_PATCHES(f, 9, L, R)
}

// Patch values only (what the function objects read every step through historyPatch: beamDisplacements.C:56-63,
// beamForcesMoments.C:56-68); the internal field stays as it was mirrored at the last write time.
void coupledTotalLagNewtonRaphsonBeamGPU::downloadPatches(const label id, volVectorField& f)
{
    List<scalar> L(3*nBeams()), R(3*nBeams());
    check(bf_get_field(ctx_, id, nullptr, L.data(), R.data()), "bf_get_field");
    UNPACK_PATCHES(f, 3, L, R)
}
void coupledTotalLagNewtonRaphsonBeamGPU::downloadPatches(const label id, surfaceVectorField& f)
{
    List<scalar> L(3*nBeams()), R(3*nBeams());
    check(bf_get_field(ctx_, id, nullptr, L.data(), R.data()), "bf_get_field");
    UNPACK_PATCHES(f, 3, L, R)
}

// The host patch-field objects stay responsible for dictionary parsing and for the time series:
// updateCoeffs() evaluates them at the new time exactly as Evolve.C:156-157 does; the values go
// to the device, which re-implements the boundary-face arithmetic (BC.C, evaluate()).
void coupledTotalLagNewtonRaphsonBeamGPU::pushBoundaryValues()
{
    volVectorField& W = solutionW();
    volVectorField& Theta = solutionTheta();
    W.boundaryFieldRef().updateCoeffs();
    Theta.boundaryFieldRef().updateCoeffs();

    for (label end = 0; end < 2; end++)
    {
        double *w = bcStage_[end][0], *th = bcStage_[end][1], *F = bcStage_[end][2], *M = bcStage_[end][3];
        for (label b = 0; b < nBeams(); b++)
        {
            const label patchI = end ? endPatchIndex(b) : startPatchIndex(b);
            const label fI = nBeams() > 1 ? 0 : b;
            const fvPatchVectorField& pW = W.boundaryField()[patchI];
            const fvPatchVectorField& pT = Theta.boundaryField()[patchI];
            vector f = vector::zero, m = vector::zero;
            if (isA<followerForceBeamDisplacementNRFvPatchVectorField>(pW))
            {
                // updateCoeffs() just set force() = Lambdaf_b & followerForce_ (no accessor for the
                // material-frame vector): recover it with the transpose of the orthogonal Lambdaf_b
                const tensor& Lfb = solutionLambda().boundaryField()[patchI][fI];
                f = (Lfb.T() & refCast<const forceBeamDisplacementNRFvPatchVectorField>(pW).force()[fI]);
            }
            else if (isA<linearSpringForceBeamDisplacementNRFvPatchVectorField>(pW))
            {
                // needs the three one-line accessors listed in INTEGRATION.md (offsetForce(),
                // springStiffness(), springDirection()) on the reference's spring patch field
                f = refCast<const linearSpringForceBeamDisplacementNRFvPatchVectorField>(pW).offsetForce()[fI];
            }
            else if (isA<forceBeamDisplacementNRFvPatchVectorField>(pW))
            {
                f = refCast<const forceBeamDisplacementNRFvPatchVectorField>(pW).force()[fI];
            }
            if (isA<momentBeamRotationNRFvPatchVectorField>(pT))
            {
                m = refCast<const momentBeamRotationNRFvPatchVectorField>(pT).moment()[fI];
            }
            for (label k = 0; k < 3; k++)
            {
                w[3*b + k] = pW[fI][k]; th[3*b + k] = pT[fI][k]; F[3*b + k] = f[k]; M[3*b + k] = m[k];
            }
        }
        // the staging arrays are page-locked and refilled only at the next time step, after evolve() has returned
        check(bf_set_bc_values_async(ctx_, end, BF_BCV_W, w), "bf_set_bc_values_async");
        check(bf_set_bc_values_async(ctx_, end, BF_BCV_THETA, th), "bf_set_bc_values_async");
        check(bf_set_bc_values_async(ctx_, end, BF_BCV_FORCE, F), "bf_set_bc_values_async");
        check(bf_set_bc_values_async(ctx_, end, BF_BCV_MOMENT, M), "bf_set_bc_values_async");
    }
}

// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

coupledTotalLagNewtonRaphsonBeamGPU::coupledTotalLagNewtonRaphsonBeamGPU(Time& runTime, const word& region)
:
    coupledTotalLagNewtonRaphsonBeam(runTime, region),
    ctx_(nullptr),
    nSeg_(nBeams(), 0),
    uploaded_(false)
{
    for (label end = 0; end < 2; end++) for (label k = 0; k < 4; k++) bcStage_[end][k] = nullptr;
    const label B = nBeams();
    List<int32_t> nSeg(B), wType(2*B), thType(2*B);
    List<scalar> length(B), CQ(3*B), CM(3*B), ARho(B), CIRho(3*B), weight(3*B), refL(9*B), sk(2*B, 0.0), sd(6*B, 0.0);
    const surfaceTensorField& refLambdaf = field<surfaceTensorField>("refLambdaf");
    const volVectorField& W = solutionW();
    const volVectorField& Theta = solutionTheta();
    const word scheme(mesh().d2dt2Schemes().found("d2dt2(W)")
        ? word(mesh().d2dt2Schemes().lookup("d2dt2(W)")) : word(mesh().d2dt2Schemes().lookup("default")));

    for (label b = 0; b < B; b++)
    {
        const labelList& cells = mesh().cellZones()[b];
        nSeg[b] = nSeg_[b] = cells.size();
        scalar len = 0;
        forAll(cells, i) { len += L()[cells[i]]; }
        length[b] = len;
        CQ[3*b] = EA(b).value(); CQ[3*b + 1] = GA(b).value(); CQ[3*b + 2] = GA(b).value();
        CM[3*b] = GJ(b).value(); CM[3*b + 1] = EIyy(b).value(); CM[3*b + 2] = EIzz(b).value();
        ARho[b] = this->ARho(b).value();
        CIRho[3*b] = kCI()*JRho(b).value(); CIRho[3*b + 1] = kCI()*IyyRho(b).value(); CIRho[3*b + 2] = kCI()*IzzRho(b).value();
        for (label k = 0; k < 3; k++) { weight[3*b + k] = rho(b).value()*A(b).value()*g().value()[k]; }
        const tensor& T = refLambdaf.boundaryField()[endPatchIndex(b)][B > 1 ? 0 : b];
        for (label k = 0; k < 9; k++) { refL[9*b + k] = T[k]; }
        for (label end = 0; end < 2; end++)
        {
            const label patchI = end ? endPatchIndex(b) : startPatchIndex(b);
            const fvPatchVectorField& pW = W.boundaryField()[patchI];
            const fvPatchVectorField& pT = Theta.boundaryField()[patchI];
            if (pW.coupled()) { notImplemented("coupled patches: shard whole beams per rank instead"); }
            label wt = BF_W_FIXED;
            if (isA<followerForceBeamDisplacementNRFvPatchVectorField>(pW)) { wt = BF_W_FOLLOWER; }
            else if (isA<linearSpringForceBeamDisplacementNRFvPatchVectorField>(pW))
            {
                wt = BF_W_SPRING;
                const auto& sp = refCast<const linearSpringForceBeamDisplacementNRFvPatchVectorField>(pW);
                sk[end*B + b] = sp.springStiffness();
                for (label k = 0; k < 3; k++) { sd[3*(end*B + b) + k] = sp.springDirection()[k]; }
            }
            else if (isA<forceBeamDisplacementNRFvPatchVectorField>(pW)) { wt = BF_W_FORCE; }
            wType[end*B + b] = wt;
            thType[end*B + b] = isA<momentBeamRotationNRFvPatchVectorField>(pT) ? BF_TH_MOMENT : BF_TH_FIXED;
        }
    }

    bf_layout lay = {B, nSeg.cdata(), length.cdata(), L().primitiveField().cdata()};
    bf_beam_props props = {CQ.cdata(), CM.cdata(), ARho.cdata(), CIRho.cdata(), weight.cdata(), refL.cdata()};
    bf_bc bc = {wType.cdata(), thType.cdata(), sk.cdata(), sd.cdata()};
    bf_options opt = {};
    opt.scheme = scheme == "Newmark" ? BF_SCHEME_NEWMARK : scheme == "Euler" ? BF_SCHEME_EULER : BF_SCHEME_STEADY;
    opt.newmarkBeta = mesh().d2dt2Schemes().lookupOrDefault<scalar>("newmarkBeta(W)", 0.25);
    opt.newmarkGamma = mesh().d2dt2Schemes().lookupOrDefault<scalar>("newmarkGamma(W)", 0.5);
    opt.device = beamProperties().lookupOrDefault<label>("device", 0);
    opt.storeIncrements = 1;
    const int rc = bf_create(&lay, &props, &bc, &opt, &ctx_);
    if (rc != BF_OK)
    {
        FatalErrorInFunction << "bf_create: " << bf_last_error(nullptr) << abort(FatalError);
    }
    if (Pstream::parRun()) { check(bf_set_reducer(ctx_, &reduceNorms, nullptr), "bf_set_reducer"); }
    for (label end = 0; end < 2; end++)
    {
        for (label k = 0; k < 4; k++)
        {
            void* ptr = nullptr;
            check(bf_host_alloc(&ptr, sizeof(double)*3*nBeams()), "bf_host_alloc");
            bcStage_[end][k] = static_cast<double*>(ptr);
        }
    }
    check(bf_set_distributed_loads
    (
        ctx_,
        reinterpret_cast<const double*>(q().primitiveField().cdata()),
        reinterpret_cast<const double*>(m().primitiveField().cdata())
    ), "bf_set_distributed_loads");

    // constant/beamMomentumContributionProperties (CTL.C:1137-1182): the base class keeps its list private, so the
    // dictionary is read again; only the coefficients cross the ABI, the per-cell arithmetic runs on the device
    IOobject mcHeader("beamMomentumContributionProperties", runTime.constant(), runTime, IOobject::MUST_READ);
    if (mcHeader.typeHeaderOk<dictionary>(true))
    {
        IOdictionary mcDict(mcHeader);
        const PtrList<entry> entries(mcDict.lookup("beamMomentumContributions"));
        List<bf_momentum_contribution> list(entries.size());
        forAll(entries, i)
        {
            const dictionary& e = entries[i].dict();
            const word type(e.lookup("beamMomentumContributionType"));
            const dictionary& k = e.subDict(type + "Coeffs");
            bf_momentum_contribution& mc = list[i];
            std::memset(&mc, 0, sizeof(mc));
            if (type == "morisonDrag")
            {
                mc.type = BF_CONTRIB_MORISON_DRAG;
                mc.coeffs[0] = readScalar(k.lookup("Cdn"));
                mc.coeffs[1] = readScalar(k.lookup("Cdt"));
                mc.coeffs[2] = readScalar(k.lookup("rhoFluid"));
            }
            else if (type == "groundContact")
            {
                mc.type = BF_CONTRIB_GROUND_CONTACT;
                mc.coeffs[0] = readScalar(k.lookup("kNormal"));
                mc.coeffs[1] = readScalar(k.lookup("cNormal"));
                mc.coeffs[2] = readScalar(k.lookup("kTangential"));
                mc.coeffs[3] = readScalar(k.lookup("muFriction"));
                mc.coeffs[4] = readScalar(k.lookup("groundZ"));
            }
            else
            {
                FatalErrorInFunction << "Unknown beamMomentumContribution type " << type << abort(FatalError);
            }
        }
        // R(bI) and the rigid translation of each beam: refW = (T & C) + translate - C  =>  translate = refW_b + C_b - T & C_b
        const volVectorField& refW = field<volVectorField>("refW");
        const volTensorField& refLambda = field<volTensorField>("refLambda");
        scalarField radius(nBeams());
        vectorField translate(nBeams(), vector::zero);
        forAll(radius, bI)
        {
            radius[bI] = R(bI);
            const label c0 = startCells()[bI];
            translate[bI] = refW[c0] + mesh().C()[c0] - (refLambda[c0] & mesh().C()[c0]);
        }
        check(bf_set_momentum_contributions
        (
            ctx_, list.size(), list.cdata(), radius.cdata(), reinterpret_cast<const double*>(translate.cdata())
        ), "bf_set_momentum_contributions");
    }
}

coupledTotalLagNewtonRaphsonBeamGPU::~coupledTotalLagNewtonRaphsonBeamGPU()
{
    bf_destroy(ctx_);
    for (label end = 0; end < 2; end++)
    {
        for (label k = 0; k < 4; k++)
        {
            if (bcStage_[end][k]) bf_host_free(bcStage_[end][k]);
        }
    }
}

// * * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * //

scalar coupledTotalLagNewtonRaphsonBeamGPU::evolve()
{
    beamModel::evolve();
    const dictionary& d = beamProperties();

    if (!uploaded_)   // state read from the time directory (restart) or the constructor defaults
    {
        upload(BF_FIELD_W, solutionW());
        upload(BF_FIELD_THETA, solutionTheta());
        upload(BF_FIELD_LAMBDA, field<volTensorField>("Lambda"));
        upload(BF_FIELD_LAMBDAF, solutionLambda());
        upload(BF_FIELD_GAMMA, field<surfaceVectorField>("Gamma"));
        upload(BF_FIELD_K, field<surfaceVectorField>("K"));
        upload(BF_FIELD_Q, field<surfaceVectorField>("Q"));
        upload(BF_FIELD_M, field<surfaceVectorField>("M"));
        if (!steadyState())
        {
            upload(BF_FIELD_U, field<volVectorField>("U"));
            upload(BF_FIELD_ACCL, field<volVectorField>("Accl"));
            upload(BF_FIELD_OMEGA, field<volVectorField>("Omega"));
            upload(BF_FIELD_DOTOMEGA, field<volVectorField>("dotOmega"));
        }
        uploaded_ = true;
    }

    // Evolve.C:61-102
    bf_tolerances tol;
    tol.nCorrectors = d.lookupOrDefault<int>("nCorrectors", 1000);
    tol.residualTol = d.lookupOrDefault<scalar>("residualTol", -1);
    if (tol.residualTol < 0)
    {
        tol.residualTol = d.found("convergenceTol") ? readScalar(d.lookup("convergenceTol")) : 1e-8;
    }
    tol.solutionTol = d.lookupOrDefault<scalar>("solutionTol", 1e-10);
    tol.absoluteTol = d.lookupOrDefault<scalar>("absoluteTol", 1e-30);
    tol.divergenceTol = d.lookupOrDefault<scalar>("divergenceTol", 1e6);

    pushBoundaryValues();
    if (pointForces().size())   // constant/pointForces, Evolve.C:211-219: the series at the new time
    {
        List<int64_t> cells(pointForces().size());
        scalarField zeta(pointForces().size());
        vectorField F0(pointForces().size());
        forAll(pointForces(), pfI)
        {
            cells[pfI] = pointForces()[pfI].first().first();
            zeta[pfI] = pointForces()[pfI].first().second();
            F0[pfI] = pointForces()[pfI].second()(runTime().value());
        }
        check
        (
            bf_set_point_forces(ctx_, cells.size(), cells.cdata(), zeta.cdata(), reinterpret_cast<const double*>(F0.cdata())),
            "bf_set_point_forces"
        );
    }
    check(bf_begin_step(ctx_, runTime().deltaTValue()), "bf_begin_step");

    bf_step_out out;
    const int rc = bf_evolve(ctx_, &tol, &out);
    iOuterCorr() = out.nIter;
    if (rc == BF_DIVERGED || rc == BF_MAXITER)   // Evolve.C:1003-1022
    {
        FatalErrorInFunction << bf_last_error(ctx_) << abort(FatalError);
    }
    check(rc, "bf_evolve");

    Info<< "    Newton iterations: " << out.nIter
        << ", residual " << out.currentResidualNorm << ", step " << out.deltaXNorm << endl;

    // Mirror back what boundary conditions, function objects and write() look up by name.  Every step: the patch values
    // (function objects with a historyPatch).  The internal fields only when something on the host reads them: at write
    // times (runTime.write(), beamFoam.C:87), or every step with `mirrorFields everyStep;` in the Coeffs dictionary (needed
    // if a function object such as beamEnergyData reads internal fields each step; bf_beam_energy replaces that one).
    const bool everyStep = d.lookupOrDefault<word>("mirrorFields", "outputTime") == "everyStep";
    if (everyStep || runTime().outputTime() || runTime().timeIndex() == runTime().startTimeIndex() + 1)
    {
        download(BF_FIELD_W, solutionW(), true);
        download(BF_FIELD_THETA, solutionTheta(), true);
        download(BF_FIELD_DW, field<volVectorField>("DW"), true);
        download(BF_FIELD_DTHETA, field<volVectorField>("DTheta"), true);
        download(BF_FIELD_LAMBDA, field<volTensorField>("Lambda"));
        download(BF_FIELD_LAMBDAF, solutionLambda());
        download(BF_FIELD_GAMMA, field<surfaceVectorField>("Gamma"));
        download(BF_FIELD_K, field<surfaceVectorField>("K"));
        download(BF_FIELD_Q, field<surfaceVectorField>("Q"));
        download(BF_FIELD_M, field<surfaceVectorField>("M"));
        if (!steadyState())
        {
            download(BF_FIELD_U, field<volVectorField>("U"), false);
            download(BF_FIELD_ACCL, field<volVectorField>("Accl"), false);
            download(BF_FIELD_OMEGA, field<volVectorField>("Omega"), false);
            download(BF_FIELD_DOTOMEGA, field<volVectorField>("dotOmega"), false);
        }
    }
    else
    {
        downloadPatches(BF_FIELD_W, solutionW());
        downloadPatches(BF_FIELD_THETA, solutionTheta());
        downloadPatches(BF_FIELD_Q, field<surfaceVectorField>("Q"));
        downloadPatches(BF_FIELD_M, field<surfaceVectorField>("M"));
    }
    solutionW().storePrevIter();
    solutionTheta().storePrevIter();
    solutionDW() = solutionW() - solutionW().oldTime();   // WIncrement_ (Evolve.C:775)

    return out.initialResidualNorm;
}

} // End namespace beamModels
} // End namespace Foam
