// beam_contrib.cuh — the run-time selectable beamMomentumContribution sources (Evolve.C:539-571) on the device.
//
// Both reference classes (morisonDragContribution.C:103-195, groundContactContribution.C:111-236) are explicit
// per-cell sources with a zero Jacobian.  They need, per cell, the mid-point derivative of the Hermite spline through
// the CURRENT beam points and tangents:
//   points    currentBeamPoints   CTL.C:1263-1427: patch face positions C_f + refW_f + W_b at the ends, and on internal
//                                 faces the Hermite mid-point between the two cell centres, with cell tangents
//                                 = normalised (upper face - lower face) of the linearly interpolated face positions
//   tangents  currentBeamTangents CTL.C:1518-1645: normalised snGrad(C + refW + W), start patch sign-flipped
//   spline    HermiteSpline.C:190-217 (chords), :34-56 calcParam, :342-386 arcLength (composite Simpson, n = 12 per
//             segment), :300-340 localParameter (secant iteration to 1e-4), :468-513 firstDerivative
// One thread per (beam, cell): the stencil is W of cells i-2..i+2 and the patch values, the arithmetic a few hundred
// FP64 operations; it runs once per Newton iteration ahead of the sweeps and leaves its result in the per-cell
// source array that the sweeps already read for the point forces (pfSrc, subtracted from the source), so a case without
// contributions carries nothing of this in the hot loop.
#pragma once
#include "beam_kernels.cuh"
#include "../../include/beamgpu.h"

struct ContribParams {
    int n;                                        // contributions in the list
    int type[BF_MAX_CONTRIBUTIONS];
    double coeffs[BF_MAX_CONTRIBUTIONS][6];
    const double* radius;                         // [Bpad] R(bI)
    const double* translation;                    // [3][Bpad]
    int keepPointForces;                          // pfSrc already holds the point-force source of this iteration
};

struct HSeg { // one spline segment: end points, unit end tangents, chord length
    V3 p0, p1, t0, t1;
    double c;
};
DEV V3 hsParamFirstDerivative(const HSeg& s, double zeta) // HermiteSpline.C:515-536
{
    const double z2 = zeta * zeta;
    const double dH0 = 3 * (z2 - 1) / 4, dH1 = -3 * (z2 - 1) / 4;
    const double dHt0 = (3 * z2 - 2 * zeta - 1) / 4, dHt1 = (3 * z2 + 2 * zeta - 1) / 4;
    return v3(s.p0.x * dH0 + s.p1.x * dH1 + s.c * (s.t0.x * dHt0 + s.t1.x * dHt1) / 2,
              s.p0.y * dH0 + s.p1.y * dH1 + s.c * (s.t0.y * dHt0 + s.t1.y * dHt1) / 2,
              s.p0.z * dH0 + s.p1.z * dH1 + s.c * (s.t0.z * dHt0 + s.t1.z * dHt1) / 2);
}
__device__ __noinline__ double hsArcLength(const HSeg& s, double zeta) // :342-386
{
    if (zeta < -1 + BF_SMALL) return 0;
    int n = (int)(12 * (zeta + 1) / 2);
    if ((n % 2) > 0) n += 1;
    if (n < 2) n = 2;
    const double h = (zeta + 1) / n;
    double arcLen = 0;
    for (int i = 1; i <= n / 2; i++)
        arcLen += h * (mag(hsParamFirstDerivative(s, -1 + (2 * i - 2) * h)) + 4 * mag(hsParamFirstDerivative(s, -1 + (2 * i - 1) * h)) +
                       mag(hsParamFirstDerivative(s, -1 + (2 * i) * h))) / 3;
    return arcLen;
}
DEV double hsLocalParameter(const HSeg& s, double arcLen, double totalArcLength) // :300-340
{
    if (arcLen > totalArcLength - BF_SMALL) return 1;
    if (arcLen < BF_SMALL) return -1;
    double residual, zeta0 = -1, zeta1 = 1, zeta2 = 0;
    int guard = 0;
    do {
        const double f0 = hsArcLength(s, zeta0) - arcLen, f1 = hsArcLength(s, zeta1) - arcLen;
        zeta2 = zeta1 - f1 * (zeta1 - zeta0) / (f1 - f0);
        residual = fabs(zeta2 - zeta1) / 2;
        zeta0 = zeta1;
        zeta1 = zeta2;
    } while (residual > 1e-4 && ++guard < 64); // the reference has no guard; 64 secant steps are never reached
    return zeta2;
}

// dRdScell of cell i of beam b, and the current cell velocity U.
template <int SCHEME>
DEV void cellSplineAndVelocity(const BeamParams& p, const ContribParams& cp, const int b, const int i, V3& dRdS, V3& U,
                               double& zCoord)
{
    const size_t T = (size_t)(b >> 5), sb = p.Bpad;
    const int lane = b & 31, n = p.nSeg[b];
    const double dZ = p.dZ[b], dcI = 1.0 / dZ, dcB = 2.0 / dZ;
    const V3 dR0 = v3(p.refL[b], p.refL[3 * sb + b], p.refL[6 * sb + b]); // T & e_x
    const V3 tr = cp.translation ? ld3(cp.translation, b, sb) : v3(0, 0, 0);
    auto Wc = [&](int k) { return ld3(p.W, tileRow(T, p.nmax, k, 3, lane), TS); };
    auto patchW = [&](int end) { // boundary value after updateCoeffs()
        const size_t i3 = (size_t)end * 3 * sb + b;
        return p.wType[end * sb + b] == BFK_W_FIXED ? ld3(p.wPresc, i3, sb) : ld3(p.Wb, i3, sb);
    };
    // C + refW = T & C + translate (setInitialPositionBeam.C:236-290), C_k = ((k + 0.5) dZ, 0, 0), Cf_j = (j dZ, 0, 0)
    auto curC = [&](int k) { return ((k + 0.5) * dZ) * dR0 + tr + Wc(k); };
    auto curCfLin = [&](int j) { // mesh.Cf() + refWf_ + fvc::interpolate(W_)   CTL.C:1281-1282
        if (j <= 0) return tr + patchW(0);
        if (j >= n) return (n * dZ) * dR0 + tr + patchW(1);
        return (j * dZ) * dR0 + tr + (0.5 * Wc(j - 1) + 0.5 * Wc(j));
    };
    auto cellTangent = [&](int k) { // :1284-1313
        const V3 v = curCfLin(k + 1) - curCfLin(k);
        return (1.0 / (mag(v) + BF_SMALL)) * v;
    };
    auto point = [&](int j) { // :1320-1331, HermiteSpline::position(p0, t0, p1, t1, 0)  HermiteSpline.C:1096-1128
        if (j <= 0 || j >= n) return curCfLin(j);
        const V3 p0 = curC(j - 1), p1 = curC(j), t0 = cellTangent(j - 1), t1 = cellTangent(j);
        const double c = mag(p1 - p0);
        return v3(p0.x * 0.5 + p1.x * 0.5 + c * (0.25 * t0.x + -0.25 * t1.x) / 2, p0.y * 0.5 + p1.y * 0.5 + c * (0.25 * t0.y + -0.25 * t1.y) / 2,
                  p0.z * 0.5 + p1.z * 0.5 + c * (0.25 * t0.z + -0.25 * t1.z) / 2);
    };
    auto tangent = [&](int j) { // :1546-1556 then HermiteSpline.C:206 (normalised once more)
        V3 v;
        if (j <= 0) v = dcB * (curCfLin(0) - curC(0));
        else if (j >= n) v = dcB * (curCfLin(n) - curC(n - 1));
        else v = dcI * (curC(j) - curC(j - 1));
        v = (1.0 / mag(v)) * v;
        if (j <= 0) v = -v;
        return (1.0 / mag(v)) * v;
    };
    HSeg s;
    s.p0 = point(i); s.p1 = point(i + 1); s.t0 = tangent(i); s.t1 = tangent(i + 1);
    s.c = mag(s.p1 - s.p0);
    // param_[i+1] - param_[i]: the reference differences the running sum; the segment length itself differs from that
    // by one rounding of the sum, which the 1e-4 tolerance of the secant iteration does not see
    const double segLen = hsArcLength(s, 1.0);
    const double zeta = hsLocalParameter(s, segLen / 2, segLen);
    if (zeta <= -1) dRdS = s.t0;
    else if (zeta >= 1.0) dRdS = s.t1;
    else {
        const V3 d = hsParamFirstDerivative(s, zeta);
        dRdS = (1.0 / mag(d)) * d;
    }
    // U as the assembly of this iteration sees it (Newmark: stored as the step constant U*, predictor on the first iteration)
    const size_t c3 = tileRow(T, p.nmax, i, 3, lane);
    U = v3(0, 0, 0);
    if (SCHEME == BFK_NEWMARK) {
        const V3 Us = ld3(p.U, c3, TS), Ao = ld3(p.Ac, c3, TS);
        if (p.first) {
            const V3 Uo = Us + p.gdtPrev * Ao;
            const V3 Ac = v3(-p.nk1 * Uo.x - p.nk2 * Ao.x, -p.nk1 * Uo.y - p.nk2 * Ao.y, -p.nk1 * Uo.z - p.nk2 * Ao.z);
            U = Uo + p.dt * ((1.0 - p.gamma) * Ao + p.gamma * Ac);
        } else
            U = Us + p.gdt * Ao;
    } else if (SCHEME == BFK_EULER)
        U = ld3(p.U, c3, TS);
    // refW.z + W.z (groundContactContribution.C:185): refW = T & C + translate - C and C.z = 0
    zCoord = ((i + 0.5) * dZ) * dR0.z + tr.z + Wc(i).z;
}

template <int SCHEME>
__global__ void beam_momentum_contributions(const __grid_constant__ BeamParams p, const ContribParams cp)
{
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = (int)(g & 31);
    const int i = (int)((g >> 5) % (size_t)p.nmax);
    const size_t T = (g >> 5) / (size_t)p.nmax;
    const int b = (int)(T * 32 + lane);
    if (b >= p.B || i >= p.nSeg[b]) return;
    V3 d, U;
    double z;
    cellSplineAndVelocity<SCHEME>(p, cp, b, i, d, U, z);
    const double R = cp.radius[b];
    const double L = p.Lc ? p.Lc[tileRow(T, p.nmax, i, 1, lane)] : p.dZ[b];
    const double Ud = dot(U, d);
    const V3 Ut = Ud * d, Un = U - Ud * d;
    const V3 UtHat = (1.0 / (mag(Ut) + BF_SMALL)) * Ut, UnHat = (1.0 / (mag(Un) + BF_SMALL)) * Un;
    V3 sum = v3(0, 0, 0);
    for (int k = 0; k < cp.n; k++) {
        const double* c = cp.coeffs[k];
        if (cp.type[k] == BF_CONTRIB_MORISON_DRAG) { // morisonDragContribution.C:103-176
            const double Fdn = c[2] * c[0] * R * L * dot(Un, Un), Fdt = c[2] * c[1] * R * L * dot(Ut, Ut);
            sum = sum + (Fdn * UnHat + Fdt * UtHat);
        } else if (z < c[4]) { // groundContactContribution.C:183-213
            const double fn = (2 * R * c[0] * (c[4] - z)) - (2 * R * c[1] * fmax(U.z, 0.0));
            const V3 ft = c[2] * 2.0 * R * mag(Ut) >= c[3] * fabs(fn) ? (-c[3] * fabs(fn)) * UtHat : (-2.0 * R * c[2]) * Ut;
            sum = sum + v3(ft.x, ft.y, fn + ft.z);
        }
    }
    // Evolve.C:553-556 adds the contribution to the source; the sweeps SUBTRACT pfSrc
    const size_t c6 = tileRow(T, p.nmax, i, 6, lane);
    const double s3[3] = {sum.x, sum.y, sum.z};
#pragma unroll
    for (int r = 0; r < 6; r++) {
        const double base = cp.keepPointForces ? p.pfSrc[c6 + (size_t)r * TS] : 0.0;
        p.pfSrc[c6 + (size_t)r * TS] = r < 3 ? base - s3[r] : base;
    }
}
