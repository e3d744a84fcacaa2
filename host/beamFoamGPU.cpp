// beamFoamGPU — a C++ host for the C ABI of include/beamgpu.h that needs no OpenFOAM.
//
// It reads the reference's own case directories (constant/beamProperties, constant/g, 0/W, 0/Theta, 0/q, 0/m,
// constant/pointForces, system/controlDict, system/fvSchemes and the (time value) tables of the boundary-condition
// series) and drives the Newton-iteration path through the ABI exactly as the reference's applications do:
//
//   beamFoamGPU createBeamMesh          -case <dir>     applications/utilities/createBeamMesh/createBeamMesh.C:100-135
//   beamFoamGPU setInitialPositionBeam  -case <dir> -cellZone beam_k -translate '(x y z)' -rotateAngle '((ax ay az) deg)'
//                                                       applications/utilities/setInitialPositionBeam/setInitialPositionBeam.C:160-305
//   beamFoamGPU beamFoam                -case <dir>     applications/solvers/beamFoam/beamFoam.C:57-94
//   beamFoamGPU Allrun                  -case <dir>     interprets the `runApplication ...` lines of the case's Allrun
//
// The implementation of the ABI is dlopen()ed: libbeamgpu.so (CUDA, sm_100a) unless -lib names another file; there is
// no fallback — a missing library or a missing device is a fatal error, as a missing solver library is in OpenFOAM.
// What stays on the host is what stays on the host in the reference: dictionaries, the BC time series
// (interpolationTable, outOfBounds clamp), the time loop and the history files.
//
// Output (instead of the function objects of system/controlDict): <case>/history/beamConvergenceData.dat
// (time, iOuterCorr, initial residual: functionObjects/beamConvergenceData) and history/beamDisplacements_right.dat
// (time, then W_b and Theta_b of the right patch of every beam: functionObjects/beamDisplacements); every writeInterval
// steps the time directory <case>/<time>/{W,Theta} in OpenFOAM ASCII format (runTime.write()).
#include "foamDictionary.hpp"
#include "../include/beamgpu.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

namespace
{
using foam::Dict;
using foam::Value;
typedef std::vector<double> vec;

[[noreturn]] void fatal(const std::string& msg)
{
    fprintf(stderr, "\n--> FOAM FATAL ERROR (beamFoamGPU):\n%s\n\n", msg.c_str());
    exit(1);
}

// ---- the ABI, resolved at run time ---------------------------------------------------------------------------------
struct Abi {
    void* h = nullptr;
#define BF_FN(name) decltype(&::name) name = nullptr;
    BF_FN(bf_create) BF_FN(bf_destroy) BF_FN(bf_last_error) BF_FN(bf_version) BF_FN(bf_backend) BF_FN(bf_num_cells)
    BF_FN(bf_set_field) BF_FN(bf_get_field) BF_FN(bf_set_distributed_loads) BF_FN(bf_set_point_forces)
    BF_FN(bf_set_bc_values) BF_FN(bf_set_momentum_contributions) BF_FN(bf_begin_step) BF_FN(bf_evolve) BF_FN(bf_launch_count)
#undef BF_FN
    void load(const std::string& path)
    {
        h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
        if (!h) fatal("cannot load the solver library " + path + ": " + dlerror() + "\n(there is no CPU fallback)");
#define BF_SYM(name)                                                                     \
    name = (decltype(name))dlsym(h, #name);                                              \
    if (!name) fatal(std::string("symbol ") + #name + " is missing from " + path);
        BF_SYM(bf_create) BF_SYM(bf_destroy) BF_SYM(bf_last_error) BF_SYM(bf_version) BF_SYM(bf_backend) BF_SYM(bf_num_cells)
        BF_SYM(bf_set_field) BF_SYM(bf_get_field) BF_SYM(bf_set_distributed_loads) BF_SYM(bf_set_point_forces)
        BF_SYM(bf_set_bc_values) BF_SYM(bf_set_momentum_contributions) BF_SYM(bf_begin_step) BF_SYM(bf_evolve) BF_SYM(bf_launch_count)
#undef BF_SYM
    }
};

// ---- interpolationTable<vector>, outOfBounds clamp (the only mode the tutorials use) ---------------------------------
struct Table {
    std::vector<double> t;
    std::vector<vec> v;
    bool empty() const { return t.empty(); }
    vec operator()(double x) const
    {
        if (x <= t.front()) return v.front();
        if (x >= t.back()) return v.back();
        size_t i = 0;
        while (i + 2 < t.size() && t[i + 1] <= x) i++;
        const double s = (x - t[i]) / (t[i + 1] - t[i]);
        vec r(v[i].size());
        for (size_t k = 0; k < r.size(); k++) r[k] = v[i][k] + s * (v[i + 1][k] - v[i][k]);
        return r;
    }
};
Table readTable(const Dict& d, const std::string& caseDir)
{
    // `forceSeries { "fileName|file" "$FOAM_CASE/constant/timeVsForce"; outOfBounds clamp; }`
    const std::string oob = d.wordOr("outOfBounds", "clamp");
    if (oob != "clamp") fatal("outOfBounds " + oob + " is not supported (the reference's cases use clamp) in " + d.file);
    const std::string path = foam::expand(d.word("file"), caseDir);
    const Value L = foam::readList(path);
    Table T;
    for (const Value& row : L.list) {
        if (row.kind != Value::LIST || row.list.size() != 2) fatal("malformed (time value) row in " + path);
        T.t.push_back(row.list[0].v);
        vec val;
        if (row.list[1].kind == Value::LIST) for (const Value& c : row.list[1].list) val.push_back(c.v);
        else val.push_back(row.list[1].v);
        T.v.push_back(val);
    }
    if (T.t.empty()) fatal("empty table " + path);
    return T;
}

// ---- the six patch-field classes: dictionary data and updateCoeffs() time lookup only ------------------------------
// (src/wireBunchingModels/beamModels/fvPatchFields/<type>/<type>FvPatchVectorField.C; the boundary-face arithmetic,
//  assembly closure and evaluate(), runs behind the ABI)
struct PatchField {
    std::string type;
    int code = 0, kind = 0;
    vec value = {0, 0, 0};      // "value uniform (..)" / "force uniform (..)" / "moment uniform (..)"
    Table series, transition;
    double springK = 0;
    vec springDir = {1, 0, 0};
    vec at(double t) const
    {
        if (!series.empty()) return series(t);
        if (!transition.empty()) { const double s = transition(t)[0]; return {s * value[0], s * value[1], s * value[2]}; }
        return value;
    }
};
vec uniformVector(const Dict& d, const std::string& key)
{
    if (!d.found(key)) return {0, 0, 0};
    const vec v = d.vector(key);
    if (v.size() != 3) fatal("keyword " + key + " is not a vector in " + d.file);
    return v;
}
PatchField readPatchField(const Dict& d, bool isW, const std::string& caseDir, const std::string& patch)
{
    PatchField p;
    p.type = d.word("type");
    auto series = [&](const char* key) { return d.found(key) ? readTable(d.subDict(key), caseDir) : Table(); };
    if (isW) {
        if (p.type == "fixedValue") { p.code = BF_W_FIXED; p.kind = BF_BCV_W; p.value = uniformVector(d, "value"); }
        else if (p.type == "fixedDisplacement") { // fixedDisplacement...C:160-198
            p.code = BF_W_FIXED; p.kind = BF_BCV_W; p.value = uniformVector(d, "value"); p.series = series("displacementSeries");
        } else if (p.type == "forceBeamDisplacementNR") { // forceBeamDisplacementNR...C:186-209
            p.code = BF_W_FORCE; p.kind = BF_BCV_FORCE; p.value = uniformVector(d, "force");
            p.series = series("forceSeries"); p.transition = series("transitionSeries");
        } else if (p.type == "followerForceBeamDisplacementNR") { // followerForce...C:183-265
            p.code = BF_W_FOLLOWER; p.kind = BF_BCV_FORCE; p.series = series("followerForceSeries");
            if (p.series.empty()) fatal("followerForceSeries is not defined for patch " + patch);
        } else if (p.type == "linearSpringForceBeamDisplacementNR") { // linearSpringForce...C:85-139,204-217
            p.code = BF_W_SPRING; p.kind = BF_BCV_FORCE;
            const Dict& c = d.subDict("linearSpringCoeffs");
            p.springK = c.scalar("springStiffness");
            p.springDir = c.vector("springDirection");
            p.series = series("offsetForceSeries");
        } else fatal("unknown patch field type " + p.type + " for W on patch " + patch);
    } else {
        if (p.type == "fixedValue") { p.code = BF_TH_FIXED; p.kind = BF_BCV_THETA; p.value = uniformVector(d, "value"); }
        else if (p.type == "fixedRotation") { // fixedRotation...C:160-213
            p.code = BF_TH_FIXED; p.kind = BF_BCV_THETA; p.value = uniformVector(d, "value"); p.series = series("thetaSeries");
        } else if (p.type == "momentBeamRotationNR") { // momentBeamRotationNR...C:168-181
            p.code = BF_TH_MOMENT; p.kind = BF_BCV_MOMENT; p.value = uniformVector(d, "moment"); p.series = series("momentSeries");
        } else fatal("unknown patch field type " + p.type + " for Theta on patch " + patch);
    }
    return p;
}

// ---- the case ---------------------------------------------------------------------------------------------------------
struct PointForce { int64_t cell; double zeta; Table series; vec constant; };
struct BeamCase {
    std::string dir;
    int nBeams = 0;
    std::vector<int32_t> nSeg;
    vec length, A, Iyy, Izz, J, E, G, rho, R;
    std::vector<PatchField> W[2], Th[2]; // [end][beam]
    int scheme = BF_SCHEME_STEADY;
    double beta = 0.25, gamma = 0.5;
    bf_tolerances tol;
    double kCI = 1.0;
    vec g = {0, 0, 0};
    bool hasQ = false, hasM = false;
    vec q = {0, 0, 0}, m = {0, 0, 0}; // uniform internalField of 0/q, 0/m
    std::vector<PointForce> pointForces;
    vec refL;                          // [nBeams*9], identity unless setInitialPositionBeam ran
    vec translation;                   // [nBeams*3]
    std::vector<bf_momentum_contribution> contributions; // constant/beamMomentumContributionProperties
    bool rotated = false;
    double deltaT = 1, endTime = 1;
    int writeInterval = 0;             // system/controlDict writeControl timeStep; writeInterval n; (0: no field output)
    int writePrecision = 6;
    int device = 0;
};
bool exists(const std::string& p) { struct stat s; return stat(p.c_str(), &s) == 0; }

// beamModel.C:400-560: the `beams ( beam_k { ... } )` list, cross-section constants, E/G/rho local or global
void readBeamProperties(BeamCase& c)
{
    const auto bp = foam::readDictionary(c.dir + "/constant/beamProperties");
    const std::string model = bp->word("beamModel");
    if (model != "coupledTotalLagNewtonRaphsonBeam" && model != "coupledTotalLagNewtonRaphsonBeamGPU")
        fatal("Unknown beamModel type " + model + "\nValid beamModel types are:\n(coupledTotalLagNewtonRaphsonBeam coupledTotalLagNewtonRaphsonBeamGPU)");
    const Dict& coeffs = bp->subDict(model + "Coeffs");
    const foam::Entry& beams = bp->lookup("beams");
    if (beams.values.empty() || beams.values[0].kind != Value::LIST) fatal("keyword beams is not a list in " + bp->file);
    const bool gE = coeffs.found("E"), gG = coeffs.found("G"), gR = coeffs.found("rho");
    for (const Value& b : beams.values[0].list) {
        if (b.kind != Value::NAMED_DICT) fatal("entries of the beams list must be dictionaries in " + bp->file);
        const Dict& d = *b.dict;
        const std::string cs = d.word("crossSectionModel");
        const Dict& cd = d.subDict(cs + "CrossSectionModelDict");
        const double sA = cd.scalarOr("scalingArea", 1.0), sI = cd.scalarOr("scalingSecondMomentArea", 1.0);
        double A, Ixx, Iyy, R;
        if (cs == "circle") { // circle.H:123-144
            const double r = cd.scalar("radius");
            A = sA * M_PI * r * r; Ixx = Iyy = sI * M_PI * r * r * r * r / 4; R = r;
        } else if (cs == "rectangle") { // rectangle.H:145-166
            const double bb = cd.scalar("b"), h = cd.scalar("h");
            A = sA * bb * h; Ixx = sI * bb * h * h * h / 12; Iyy = sI * h * bb * bb * bb / 12; R = h / 2; // rectangle.H:133-136
        } else fatal("Unknown crossSectionModel " + cs + " (circle, rectangle are supported)");
        c.A.push_back(A);
        c.R.push_back(R);
        c.Iyy.push_back(Ixx); // beamModel.H:449-489: Iyy(b) = crossSection.Ixx(), Izz(b) = crossSection.Iyy(), J = IT()
        c.Izz.push_back(Iyy);
        c.J.push_back(cs == "circle" ? 2 * Ixx : Ixx + Iyy);
        c.length.push_back(d.scalar("length"));
        c.nSeg.push_back((int32_t)d.scalar("nSegments"));
        const bool local = d.found("E") || d.found("G") || d.found("rho");
        if (local) {
            if (gE || gG || gR) fatal("Both local and global inputs for beam mechanical properties found!");
            c.E.push_back(d.scalar("E")); c.G.push_back(d.scalar("G")); c.rho.push_back(d.scalarOr("rho", 0.0));
        } else if (gE || gG || gR) {
            c.E.push_back(coeffs.scalarOr("E", 0.0)); c.G.push_back(coeffs.scalarOr("G", 0.0)); c.rho.push_back(coeffs.scalarOr("rho", 0.0));
        } else fatal("Missing E, G: neither per-beam 'E'/'G' nor global 'E'/'G' in " + bp->file);
    }
    c.nBeams = (int)c.nSeg.size();
    if (!c.nBeams) fatal("no beams in " + bp->file);
    // Evolve.C:61-102
    c.tol.nCorrectors = (int32_t)coeffs.scalarOr("nCorrectors", 1000);
    c.tol.residualTol = coeffs.found("residualTol") ? coeffs.scalar("residualTol") : coeffs.scalarOr("convergenceTol", 1e-8);
    c.tol.solutionTol = coeffs.scalarOr("solutionTol", 1e-10);
    c.tol.absoluteTol = coeffs.scalarOr("absoluteTol", 1e-30);
    c.tol.divergenceTol = coeffs.scalarOr("divergenceTol", 1e6);
    c.kCI = coeffs.scalarOr("scalingInertiaTensor", 1.0);
    c.device = (int)coeffs.scalarOr("device", 0);
}

std::string patchName(const BeamCase& c, int end, int b, const Dict& boundary)
{
    // createBeamMesh.C:440-520: left_k / right_k for a bundle, left / right for a single beam
    const std::string base = end ? "right" : "left", k = base + "_" + std::to_string(b);
    if (boundary.found(k)) return k;
    if (c.nBeams == 1 && boundary.found(base)) return base;
    fatal("patch " + k + " is missing from the boundaryField of " + boundary.file);
}

void readCase(BeamCase& c)
{
    readBeamProperties(c);
    // system/fvSchemes: CTL.C:668-765
    {
        const auto fs = foam::readDictionary(c.dir + "/system/fvSchemes");
        const Dict& d2 = fs->subDict("d2dt2Schemes");
        const std::string s = d2.word("default");
        if (s == "steadyState") c.scheme = BF_SCHEME_STEADY;
        else if (s == "Newmark") c.scheme = BF_SCHEME_NEWMARK;
        else if (s == "Euler") c.scheme = BF_SCHEME_EULER;
        else fatal("Please specify the d2dt2Schemes: steadyState, Euler, Newmark (found " + s + ")");
        c.beta = d2.scalarOr("newmarkBeta(W)", 0.25);
        c.gamma = d2.scalarOr("newmarkGamma(W)", 0.5);
    }
    if (c.scheme != BF_SCHEME_STEADY)
        for (double r : c.rho)
            if (r == 0.0) fatal("The time integration scheme provided is transient but density 'rho' is not specified!!"); // Evolve.C:323-335
    {
        const auto cd = foam::readDictionary(c.dir + "/system/controlDict");
        c.deltaT = cd->scalar("deltaT");
        if (cd->wordOr("writeControl", "timeStep") == "timeStep") c.writeInterval = (int)cd->scalarOr("writeInterval", 0);
        c.writePrecision = (int)cd->scalarOr("writePrecision", 6);
        c.endTime = cd->scalar("endTime");
    }
    if (exists(c.dir + "/constant/g")) c.g = foam::readDictionary(c.dir + "/constant/g")->vector("value");
    for (int isW = 1; isW >= 0; isW--) {
        const std::string f = c.dir + (isW ? "/0/W" : "/0/Theta");
        const auto fd = foam::readDictionary(f);
        const Dict& bf = fd->subDict("boundaryField");
        for (int end = 0; end < 2; end++)
            for (int b = 0; b < c.nBeams; b++) {
                const std::string pn = patchName(c, end, b, bf);
                (isW ? c.W : c.Th)[end].push_back(readPatchField(bf.subDict(pn), isW, c.dir, pn));
            }
    }
    for (int isQ = 1; isQ >= 0; isQ--) {
        const std::string f = c.dir + (isQ ? "/0/q" : "/0/m");
        if (!exists(f)) continue;
        const auto fd = foam::readDictionary(f);
        const foam::Entry& e = fd->lookup("internalField");
        if (e.values.empty() || e.values[0].s != "uniform") fatal("only a uniform internalField is supported in " + f);
        (isQ ? c.q : c.m) = fd->vector("internalField");
        (isQ ? c.hasQ : c.hasM) = true;
    }
    if (exists(c.dir + "/constant/pointForces")) { // beamModel.C:863-866: List<Tuple2<Tuple2<label, scalar>, interpolationTable<vector>>>
        const Value L = foam::readList(c.dir + "/constant/pointForces");
        for (const Value& row : L.list) {
            if (row.kind != Value::LIST || row.list.size() < 2 || row.list[0].kind != Value::LIST) fatal("malformed entry in constant/pointForces");
            PointForce pf;
            pf.cell = (int64_t)row.list[0].list[0].v;
            pf.zeta = row.list[0].list[1].v;
            const Value& tab = row.list[1];
            for (const Value& r : tab.list) {
                if (r.kind == Value::LIST && r.list.size() == 2 && r.list[1].kind == Value::LIST) {
                    pf.series.t.push_back(r.list[0].v);
                    vec v;
                    for (const Value& x : r.list[1].list) v.push_back(x.v);
                    pf.series.v.push_back(v);
                }
            }
            if (pf.series.empty()) fatal("point force without a (time force) table in constant/pointForces");
            c.pointForces.push_back(pf);
        }
    }
    // what setInitialPositionBeam left behind
    c.refL.assign((size_t)c.nBeams * 9, 0.0);
    for (int b = 0; b < c.nBeams; b++) c.refL[9 * b] = c.refL[9 * b + 4] = c.refL[9 * b + 8] = 1.0;
    c.translation.assign((size_t)c.nBeams * 3, 0.0);
    const std::string ip = c.dir + "/constant/initialPositionBeam";
    if (exists(ip)) {
        const auto d_ = foam::readDictionary(ip);
        for (int b = 0; b < c.nBeams; b++) {
            const std::string z = "beam_" + std::to_string(b);
            if (!d_->found(z)) continue;
            const vec T = d_->subDict(z).vector("rotation");
            if (T.size() != 9) fatal("rotation of " + z + " is not a tensor in " + ip);
            for (int k = 0; k < 9; k++) c.refL[9 * b + k] = T[k];
            const vec d = d_->subDict(z).vector("translation");
            for (int k = 0; k < 3; k++) c.translation[3 * b + k] = d[k];
            c.rotated = true;
        }
    }
}

// constant/beamMomentumContributionProperties (CTL.C:1137-1182): beamMomentumContributions ( name { beamMomentumContributionType T; TCoeffs {..} } )
void readMomentumContributions(BeamCase& c)
{
    const std::string f = c.dir + "/constant/beamMomentumContributionProperties";
    if (!exists(f)) return;
    const auto d = foam::readDictionary(f);
    const foam::Entry& e = d->lookup("beamMomentumContributions");
    if (e.values.empty() || e.values[0].kind != Value::LIST) fatal("keyword beamMomentumContributions is not a list in " + f);
    for (const Value& v : e.values[0].list) {
        if (v.kind != Value::NAMED_DICT) fatal("entries of beamMomentumContributions must be dictionaries in " + f);
        const std::string type = v.dict->word("beamMomentumContributionType");
        const Dict& k = v.dict->subDict(type + "Coeffs");
        bf_momentum_contribution mc;
        memset(&mc, 0, sizeof mc);
        if (type == "morisonDrag") { // morisonDragContribution.C:48-75
            mc.type = BF_CONTRIB_MORISON_DRAG;
            mc.coeffs[0] = k.scalar("Cdn"); mc.coeffs[1] = k.scalar("Cdt"); mc.coeffs[2] = k.scalar("rhoFluid");
        } else if (type == "groundContact") { // groundContactContribution.C:48-83
            mc.type = BF_CONTRIB_GROUND_CONTACT;
            mc.coeffs[0] = k.scalar("kNormal"); mc.coeffs[1] = k.scalar("cNormal"); mc.coeffs[2] = k.scalar("kTangential");
            mc.coeffs[3] = k.scalar("muFriction"); mc.coeffs[4] = k.scalar("groundZ");
        } else fatal("Unknown beamMomentumContribution type " + type + "\n\nValid beamMomentumContribution types are :\n(groundContact morisonDrag)");
        c.contributions.push_back(mc);
    }
}

// ---- createBeamMesh -------------------------------------------------------------------------------------------------
int createBeamMesh(BeamCase& c)
{
    readBeamProperties(c);
    int64_t cells = 0;
    for (int n : c.nSeg) cells += n;
    printf("createBeamMesh: %d beam(s), %lld cells, %lld internal faces, %d boundary faces\n", c.nBeams, (long long)cells,
           (long long)(cells - c.nBeams), 2 * c.nBeams);
    for (int b = 0; b < c.nBeams && b < 8; b++)
        printf("    beam_%d: cells %d, length %g, patches %s %s\n", b, c.nSeg[b], c.length[b],
               c.nBeams == 1 ? "left" : ("left_" + std::to_string(b)).c_str(), c.nBeams == 1 ? "right" : ("right_" + std::to_string(b)).c_str());
    remove((c.dir + "/constant/initialPositionBeam").c_str()); // a new mesh starts unrotated
    return 0;
}

// ---- setInitialPositionBeam -----------------------------------------------------------------------------------------
// Parses "(x y z)" or "((x y z) a)" into numbers, in order.
vec numbersOf(const std::string& s)
{
    vec r;
    const char* p = s.c_str();
    while (*p) {
        if (*p == '(' || *p == ')' || isspace((unsigned char)*p)) { p++; continue; }
        char* e;
        const double d = strtod(p, &e);
        if (e == p) fatal("cannot parse '" + s + "'");
        r.push_back(d);
        p = e;
    }
    return r;
}
int setInitialPositionBeam(BeamCase& c, const std::string& zone, const std::string& translate, const std::string& rotateAngle)
{
    if (zone.empty()) fatal("Option cellZone is not specified"); // setInitialPositionBeam.C:176-180
    vec T = {1, 0, 0, 0, 1, 0, 0, 0, 1}, d = {0, 0, 0};
    if (!translate.empty()) { d = numbersOf(translate); if (d.size() != 3) fatal("-translate needs a vector"); }
    if (!rotateAngle.empty()) { // coordinateRotations::axisAngle::rotation(axis, angle, degrees = true): Rodrigues
        const vec a = numbersOf(rotateAngle);
        if (a.size() != 4) fatal("-rotateAngle needs ((ax ay az) angle)");
        const double n = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
        const double x = a[0] / n, y = a[1] / n, z = a[2] / n, th = a[3] * M_PI / 180.0, s = std::sin(th), k = 1.0 - std::cos(th);
        T = {1 + k * (-z * z - y * y), -s * z + k * x * y, s * y + k * x * z,
             s * z + k * x * y, 1 + k * (-z * z - x * x), -s * x + k * y * z,
             -s * y + k * x * z, s * x + k * y * z, 1 + k * (-y * y - x * x)};
    }
    // keep the other zones' entries
    const std::string ip = c.dir + "/constant/initialPositionBeam";
    std::map<std::string, std::pair<vec, vec>> all;
    if (exists(ip)) {
        const auto old = foam::readDictionary(ip);
        for (const foam::Entry& e : old->entries)
            if (e.dict) all[e.key] = {e.dict->vector("translation"), e.dict->vector("rotation")};
    }
    all[zone] = {d, T};
    FILE* f = fopen(ip.c_str(), "w");
    if (!f) fatal("cannot write " + ip);
    fprintf(f, "// written by beamFoamGPU setInitialPositionBeam: refLambda = refLambdaf = rotation, refTangent = rotation & x\n");
    for (const auto& kv : all) {
        fprintf(f, "%s\n{\n    translation (%.17g %.17g %.17g);\n    rotation (", kv.first.c_str(), kv.second.first[0], kv.second.first[1], kv.second.first[2]);
        for (int k = 0; k < 9; k++) fprintf(f, "%s%.17g", k ? " " : "", kv.second.second[k]);
        fprintf(f, ");\n}\n");
    }
    fclose(f);
    printf("setInitialPositionBeam: zone %s, T = (%g %g %g %g %g %g %g %g %g)\n", zone.c_str(), T[0], T[1], T[2], T[3], T[4], T[5], T[6], T[7], T[8]);
    return 0;
}

// ---- beamFoam ---------------------------------------------------------------------------------------------------------
#define CHECK(call)                                                                                       \
    do {                                                                                                  \
        const int rc_ = (call);                                                                           \
        if (rc_ != BF_OK) fatal(std::string(#call) + " failed (" + std::to_string(rc_) + "): " + abi.bf_last_error(ctx)); \
    } while (0)

// runTime.write() (beamFoam.C:87) for the two solution fields: <case>/<time>/W and Theta as OpenFOAM ASCII volVectorFields
// (nonuniform internalField in createBeamMesh cell order, one value per end patch), readable by the reference's tools.
std::string timeName(double t)
{
    char buf[64];
    snprintf(buf, sizeof buf, "%.6g", t); // timeFormat general, timePrecision 6
    return buf;
}
void writeVolVectorField(const BeamCase& c, const std::string& dir, const char* name, const char* dims, const vec& I,
                         const vec& L, const vec& R, const std::vector<PatchField> bc[2])
{
    FILE* f = fopen((dir + "/" + name).c_str(), "w");
    if (!f) fatal("cannot write " + dir + "/" + name);
    const int P = c.writePrecision;
    fprintf(f, "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       volVectorField;\n    object      %s;\n}\n\n", name);
    fprintf(f, "dimensions      %s;\n\ninternalField   nonuniform List<vector>\n%zu\n(\n", dims, I.size() / 3);
    for (size_t i = 0; i < I.size(); i += 3) fprintf(f, "(%.*g %.*g %.*g)\n", P, I[i], P, I[i + 1], P, I[i + 2]);
    fprintf(f, ")\n;\n\nboundaryField\n{\n");
    for (int b = 0; b < c.nBeams; b++)
        for (int end = 0; end < 2; end++) {
            const std::string pn = c.nBeams == 1 ? (end ? "right" : "left") : (end ? "right_" : "left_") + std::to_string(b);
            const vec& v = end ? R : L;
            fprintf(f, "    %s\n    {\n        type            %s;\n        value           uniform (%.*g %.*g %.*g);\n    }\n", pn.c_str(),
                    bc[end][b].type.c_str(), P, v[3 * b], P, v[3 * b + 1], P, v[3 * b + 2]);
        }
    fprintf(f, "    beam\n    {\n        type            empty;\n    }\n}\n");
    fclose(f);
}

int beamFoam(BeamCase& c, const std::string& libPath, int maxSteps)
{
    readCase(c);
    readMomentumContributions(c);
    Abi abi;
    abi.load(libPath);
    const int B = c.nBeams;
    // CTL.C:538-585 (CQ_, CM_), :621-667 (ARho_, CIRho_), Evolve.C:201-209 (self-weight)
    vec CQ(3 * B), CM(3 * B), ARho(B), CI(3 * B), weight(3 * B);
    for (int b = 0; b < B; b++) {
        CQ[3 * b] = c.E[b] * c.A[b]; CQ[3 * b + 1] = CQ[3 * b + 2] = c.G[b] * c.A[b];
        CM[3 * b] = c.G[b] * c.J[b]; CM[3 * b + 1] = c.E[b] * c.Iyy[b]; CM[3 * b + 2] = c.E[b] * c.Izz[b];
        ARho[b] = c.A[b] * c.rho[b];
        CI[3 * b] = c.kCI * (c.rho[b] * c.J[b]); CI[3 * b + 1] = c.kCI * (c.rho[b] * c.Iyy[b]); CI[3 * b + 2] = c.kCI * (c.rho[b] * c.Izz[b]);
        for (int k = 0; k < 3; k++) weight[3 * b + k] = (c.rho[b] * c.A[b]) * c.g[k];
    }
    std::vector<int32_t> wType(2 * B), thType(2 * B);
    vec springK(2 * B, 0.0), springDir(6 * B, 0.0);
    for (int end = 0; end < 2; end++)
        for (int b = 0; b < B; b++) {
            wType[end * B + b] = c.W[end][b].code;
            thType[end * B + b] = c.Th[end][b].code;
            if (c.W[end][b].code == BF_W_SPRING) {
                springK[end * B + b] = c.W[end][b].springK;
                for (int k = 0; k < 3; k++) springDir[3 * (end * B + b) + k] = c.W[end][b].springDir[k];
            }
        }
    bf_layout lay = {B, c.nSeg.data(), c.length.data(), nullptr};
    bf_beam_props props = {CQ.data(), CM.data(), ARho.data(), CI.data(), weight.data(), c.rotated ? c.refL.data() : nullptr};
    bf_bc bc = {wType.data(), thType.data(), springK.data(), springDir.data()};
    bf_options opt;
    memset(&opt, 0, sizeof opt);
    opt.scheme = c.scheme; opt.newmarkBeta = c.beta; opt.newmarkGamma = c.gamma; opt.device = c.device; opt.storeIncrements = 1;
    bf_ctx* ctx = nullptr;
    if (abi.bf_create(&lay, &props, &bc, &opt, &ctx) != BF_OK) fatal(std::string("bf_create failed: ") + abi.bf_last_error(nullptr));
    printf("beamFoamGPU: %s (%s), %d beam(s), %lld cells\n", abi.bf_version(), abi.bf_backend(), B, (long long)abi.bf_num_cells(ctx));

    // "value uniform" of the fixedValue patches of 0/W, 0/Theta
    for (int isW = 1; isW >= 0; isW--) {
        vec lr[2] = {vec(3 * B, 0.0), vec(3 * B, 0.0)};
        bool any = false;
        for (int end = 0; end < 2; end++)
            for (int b = 0; b < B; b++) {
                const PatchField& p = (isW ? c.W : c.Th)[end][b];
                if (p.code != (isW ? BF_W_FIXED : BF_TH_FIXED)) continue;
                const vec v = p.at(0.0);
                for (int k = 0; k < 3; k++) { lr[end][3 * b + k] = v[k]; any = any || v[k] != 0.0; }
            }
        if (any) CHECK(abi.bf_set_field(ctx, isW ? BF_FIELD_W : BF_FIELD_THETA, nullptr, lr[0].data(), lr[1].data()));
    }
    if (!c.contributions.empty())
        CHECK(abi.bf_set_momentum_contributions(ctx, (int32_t)c.contributions.size(), c.contributions.data(), c.R.data(), c.translation.data()));
    const int64_t nCells = abi.bf_num_cells(ctx);
    if (c.hasQ || c.hasM) {
        vec q, m;
        if (c.hasQ) { q.resize(3 * nCells); for (int64_t i = 0; i < nCells; i++) for (int k = 0; k < 3; k++) q[3 * i + k] = c.q[k]; }
        if (c.hasM) { m.resize(3 * nCells); for (int64_t i = 0; i < nCells; i++) for (int k = 0; k < 3; k++) m[3 * i + k] = c.m[k]; }
        CHECK(abi.bf_set_distributed_loads(ctx, c.hasQ ? q.data() : nullptr, c.hasM ? m.data() : nullptr));
    }

    mkdir((c.dir + "/history").c_str(), 0755);
    FILE* fc = fopen((c.dir + "/history/beamConvergenceData.dat").c_str(), "w");
    FILE* fd = fopen((c.dir + "/history/beamDisplacements_right.dat").c_str(), "w");
    if (!fc || !fd) fatal("cannot write into " + c.dir + "/history");
    fprintf(fc, "# time iOuterCorr initialResidual\n");
    fprintf(fd, "# time, then per beam: Wx Wy Wz Thetax Thetay Thetaz of the right patch\n");

    // beamFoam.C:73-89: while (runTime.loop()) { beam->evolve(); beam->updateTotalFields(); runTime.write(); }
    vec vals(3 * B), Wr(3 * B), Tr(3 * B);
    long totalIter = 0;
    int step = 0;
    for (;;) {
        if (step * c.deltaT > c.endTime - 0.5 * c.deltaT) break;
        if (maxSteps >= 0 && step >= maxSteps) break;
        step++;
        const double t = step * c.deltaT;
        // W_/Theta_.boundaryFieldRef().updateCoeffs() (Evolve.C:156-157): the series at the new time
        for (int end = 0; end < 2; end++)
            for (int isW = 1; isW >= 0; isW--) {
                const std::vector<PatchField>& P = (isW ? c.W : c.Th)[end];
                bool kinds[4] = {false, false, false, false};
                for (int b = 0; b < B; b++) kinds[P[b].kind] = true;
                for (int kind = 0; kind < 4; kind++) {
                    if (!kinds[kind]) continue;
                    for (int b = 0; b < B; b++) {
                        const vec v = P[b].kind == kind ? P[b].at(t) : vec{0, 0, 0};
                        for (int k = 0; k < 3; k++) vals[3 * b + k] = v[k];
                    }
                    CHECK(abi.bf_set_bc_values(ctx, end, kind, vals.data()));
                }
            }
        if (!c.pointForces.empty()) { // pointForces()[pfI].second()(runTime().value()) (Evolve.C:219)
            std::vector<int64_t> cells;
            vec zeta, force;
            for (const PointForce& pf : c.pointForces) {
                cells.push_back(pf.cell); zeta.push_back(pf.zeta);
                const vec f = pf.series(t);
                force.insert(force.end(), f.begin(), f.end());
            }
            CHECK(abi.bf_set_point_forces(ctx, (int64_t)cells.size(), cells.data(), zeta.data(), force.data()));
        }
        CHECK(abi.bf_begin_step(ctx, c.deltaT));
        bf_step_out out;
        const int rc = abi.bf_evolve(ctx, &c.tol, &out);
        if (rc != BF_OK) fatal(std::string("evolve() at time ") + std::to_string(t) + ": " + abi.bf_last_error(ctx)); // Evolve.C:1003-1022
        totalIter += out.nIter;
        CHECK(abi.bf_get_field(ctx, BF_FIELD_W, nullptr, nullptr, Wr.data()));
        CHECK(abi.bf_get_field(ctx, BF_FIELD_THETA, nullptr, nullptr, Tr.data()));
        if (c.writeInterval > 0 && step % c.writeInterval == 0) { // runTime.write()
            vec WI(3 * nCells), TI(3 * nCells), WL(3 * B), TL(3 * B);
            CHECK(abi.bf_get_field(ctx, BF_FIELD_W, WI.data(), WL.data(), Wr.data()));
            CHECK(abi.bf_get_field(ctx, BF_FIELD_THETA, TI.data(), TL.data(), Tr.data()));
            const std::string td = c.dir + "/" + timeName(t);
            mkdir(td.c_str(), 0755);
            writeVolVectorField(c, td, "W", "[0 1 0 0 0 0 0]", WI, WL, Wr, c.W);
            writeVolVectorField(c, td, "Theta", "[0 0 0 0 0 0 0]", TI, TL, Tr, c.Th);
        }
        fprintf(fc, "%.12g %d %.17g\n", t, out.nIter, out.initialResidualNorm);
        fprintf(fd, "%.12g", t);
        for (int b = 0; b < B; b++)
            fprintf(fd, " %.17g %.17g %.17g %.17g %.17g %.17g", Wr[3 * b], Wr[3 * b + 1], Wr[3 * b + 2], Tr[3 * b], Tr[3 * b + 1], Tr[3 * b + 2]);
        fprintf(fd, "\n");
        if (step <= 3 || step % 100 == 0)
            printf("Time = %g  iOuterCorr %d  initial residual %.6e  tip W (%.9g %.9g %.9g)\n", t, out.nIter, out.initialResidualNorm,
                   Wr[3 * (B - 1)], Wr[3 * (B - 1) + 1], Wr[3 * (B - 1) + 2]);
    }
    fclose(fc);
    fclose(fd);
    printf("End: %d steps, %ld Newton iterations, %lld kernel launches\n", step, totalIter, (long long)abi.bf_launch_count(ctx));
    abi.bf_destroy(ctx);
    return 0;
}

// ---- Allrun -----------------------------------------------------------------------------------------------------------
std::vector<std::string> shellWords(const std::string& line)
{
    std::vector<std::string> w;
    std::string cur;
    bool in = false, quoted = false;
    for (char ch : line) {
        if (quoted) { if (ch == '\'') quoted = false; else cur += ch; continue; }
        if (ch == '\'') { quoted = true; in = true; continue; }
        if (isspace((unsigned char)ch)) { if (in) { w.push_back(cur); cur.clear(); in = false; } continue; }
        if (ch == '#' && !in) break;
        cur += ch; in = true;
    }
    if (in) w.push_back(cur);
    return w;
}
} // namespace

int main(int argc, char** argv)
{
    std::string app, caseDir = ".", lib, zone, translate, rotateAngle;
    int maxSteps = -1;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) fatal("option " + a + " needs a value"); return std::string(argv[++i]); };
        if (a == "-case") caseDir = next();
        else if (a == "-lib") lib = next();
        else if (a == "-cellZone") zone = next();
        else if (a == "-translate") translate = next();
        else if (a == "-rotateAngle") rotateAngle = next();
        else if (a == "-steps") maxSteps = atoi(next().c_str());
        else if (a[0] != '-' && app.empty()) app = a;
        else fatal("unknown argument " + a);
    }
    if (lib.empty()) { // $BEAMGPU_LIB, else <repo>/beamfoam_b200/libbeamgpu.so next to this executable's directory
        const char* e = getenv("BEAMGPU_LIB");
        if (e) lib = e;
        else {
            char exe[4096];
            const ssize_t n = readlink("/proc/self/exe", exe, sizeof exe - 1);
            std::string dir = n > 0 ? std::string(exe, (size_t)n) : std::string(argv[0]);
            const size_t slash = dir.rfind('/');
            dir = slash == std::string::npos ? "." : dir.substr(0, slash);
            lib = dir + "/../beamfoam_b200/libbeamgpu.so";
        }
    }
    try {
        BeamCase c;
        c.dir = caseDir;
        if (app == "createBeamMesh") return createBeamMesh(c);
        if (app == "setInitialPositionBeam") { readBeamProperties(c); return setInitialPositionBeam(c, zone, translate, rotateAngle); }
        if (app == "beamFoam") return beamFoam(c, lib, maxSteps);
        if (app == "Allrun") {
            std::ifstream f(caseDir + "/Allrun");
            if (!f) fatal("no Allrun in " + caseDir);
            std::string line;
            while (std::getline(f, line)) {
                std::vector<std::string> w = shellWords(line);
                if (w.empty() || w[0] != "runApplication") continue;
                w.erase(w.begin());
                if (w.size() >= 2 && w[0] == "-s") w.erase(w.begin(), w.begin() + 2);
                if (w.empty()) continue;
                BeamCase k;
                k.dir = caseDir;
                std::string z, tr, ra;
                for (size_t j = 1; j + 1 < w.size(); j++) {
                    if (w[j] == "-cellZone") z = w[j + 1];
                    if (w[j] == "-translate") tr = w[j + 1];
                    if (w[j] == "-rotateAngle") ra = w[j + 1];
                }
                if (w[0] == "createBeamMesh") createBeamMesh(k);
                else if (w[0] == "setInitialPositionBeam") { readBeamProperties(k); setInitialPositionBeam(k, z, tr, ra); }
                else if (w[0] == "$application" || w[0] == "beamFoam") beamFoam(k, lib, maxSteps);
                else printf("Allrun: skipping %s\n", w[0].c_str());
            }
            return 0;
        }
        fatal("usage: beamFoamGPU <createBeamMesh|setInitialPositionBeam|beamFoam|Allrun> -case <dir> [-lib <libbeamgpu.so>] [-steps n]");
    } catch (const std::exception& e) {
        fatal(e.what());
    }
}
