// beam_math.cuh — register-resident 3x3 / 6x6 FP64 helpers for the beam kernels.
//
// Device counterparts of the reference's numerics helpers:
//   spinTensor(v), axialVector      src/wireBunchingModels/numerics/spinTensor/spinTensor.C:154-168,290-299
//   rotationMatrix(theta)           src/wireBunchingModels/numerics/pseudoVector/pseudoVector.C:250-264
//   inv(tensor)                     OpenFOAM Tensor::inv (adjugate / determinant)
// Everything is fully unrolled with compile-time indices so that the small
// matrices live in registers.  Tensors are row-major (xx xy xz yx yy yz zx zy zz).
#pragma once
#include <math.h>

#define BF_SMALL 1e-15  // OpenFOAM SMALL (double)

#define DEV __device__ __forceinline__

struct V3 {
    double x, y, z;
};
struct M3 {
    double a[9];
};

DEV V3 v3(double x, double y, double z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
DEV V3 operator*(double s, V3 a) { return v3(s * a.x, s * a.y, s * a.z); }
DEV double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
DEV double mag(V3 a) { return sqrt(dot(a, a)); }

DEV M3 m3zero()
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = 0.0;
    return r;
}
DEV M3 m3eye()
{
    M3 r = m3zero();
    r.a[0] = r.a[4] = r.a[8] = 1.0;
    return r;
}
DEV M3 operator+(const M3& A, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = A.a[i] + B.a[i];
    return r;
}
DEV M3 operator-(const M3& A, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = A.a[i] - B.a[i];
    return r;
}
DEV M3 operator*(double s, const M3& A)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 9; i++) r.a[i] = s * A.a[i];
    return r;
}
DEV M3 transpose(const M3& A)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++) r.a[3 * i + j] = A.a[3 * j + i];
    return r;
}
// A & B
DEV M3 mul(const M3& A, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = A.a[3 * i] * B.a[j] + A.a[3 * i + 1] * B.a[3 + j] + A.a[3 * i + 2] * B.a[6 + j];
    return r;
}
// A & v
DEV V3 mul(const M3& A, V3 v)
{
    return v3(A.a[0] * v.x + A.a[1] * v.y + A.a[2] * v.z, A.a[3] * v.x + A.a[4] * v.y + A.a[5] * v.z,
              A.a[6] * v.x + A.a[7] * v.y + A.a[8] * v.z);
}
// c - A & v as three fused chains (no separate multiply / subtract)
DEV V3 subMul(V3 c, const M3& A, V3 v)
{
    return v3(fma(-A.a[2], v.z, fma(-A.a[1], v.y, fma(-A.a[0], v.x, c.x))),
              fma(-A.a[5], v.z, fma(-A.a[4], v.y, fma(-A.a[3], v.x, c.y))),
              fma(-A.a[8], v.z, fma(-A.a[7], v.y, fma(-A.a[6], v.x, c.z))));
}
// A.T() & v - c as three fused chains
DEV V3 mulTSub(const M3& A, V3 v, V3 c)
{
    return v3(fma(A.a[6], v.z, fma(A.a[3], v.y, fma(A.a[0], v.x, -c.x))),
              fma(A.a[7], v.z, fma(A.a[4], v.y, fma(A.a[1], v.x, -c.y))),
              fma(A.a[8], v.z, fma(A.a[5], v.y, fma(A.a[2], v.x, -c.z))));
}
// E - C & B as nine fused chains
DEV M3 subMulM(const M3& E, const M3& C, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = fma(-C.a[3 * i + 2], B.a[6 + j], fma(-C.a[3 * i + 1], B.a[3 + j], fma(-C.a[3 * i], B.a[j], E.a[3 * i + j])));
    return r;
}
// A.T() & v
DEV V3 mulT(const M3& A, V3 v)
{
    return v3(A.a[0] * v.x + A.a[3] * v.y + A.a[6] * v.z, A.a[1] * v.x + A.a[4] * v.y + A.a[7] * v.z,
              A.a[2] * v.x + A.a[5] * v.y + A.a[8] * v.z);
}
// spinTensor(v): S(v) w = v x w
DEV M3 spin(V3 v)
{
    M3 r;
    r.a[0] = 0.0;  r.a[1] = -v.z; r.a[2] = v.y;
    r.a[3] = v.z;  r.a[4] = 0.0;  r.a[5] = -v.x;
    r.a[6] = -v.y; r.a[7] = v.x;  r.a[8] = 0.0;
    return r;
}
// A & spinTensor(t): row r of the result is (row r of A) x t
DEV M3 mulSpinR(const M3& A, V3 t)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const double ax = A.a[3 * i], ay = A.a[3 * i + 1], az = A.a[3 * i + 2];
        r.a[3 * i] = ay * t.z - az * t.y;
        r.a[3 * i + 1] = az * t.x - ax * t.z;
        r.a[3 * i + 2] = ax * t.y - ay * t.x;
    }
    return r;
}
// spinTensor(t) & A: column c of the result is t x (column c of A)
DEV M3 mulSpinL(V3 t, const M3& A)
{
    M3 r;
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const double ax = A.a[c], ay = A.a[3 + c], az = A.a[6 + c];
        r.a[c] = t.y * az - t.z * ay;
        r.a[3 + c] = t.z * ax - t.x * az;
        r.a[6 + c] = t.x * ay - t.y * ax;
    }
    return r;
}
// Lam & (diag(c) & Lam.T())
DEV M3 conjDiag(const M3& Lm, V3 c)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = Lm.a[3 * i] * (c.x * Lm.a[3 * j]) + Lm.a[3 * i + 1] * (c.y * Lm.a[3 * j + 1]) +
                             Lm.a[3 * i + 2] * (c.z * Lm.a[3 * j + 2]);
    return r;
}
// diag(c) & Lam.T()
DEV M3 diagMulT(V3 c, const M3& Lm)
{
    M3 r;
#pragma unroll
    for (int j = 0; j < 3; j++) {
        r.a[j] = c.x * Lm.a[3 * j];
        r.a[3 + j] = c.y * Lm.a[3 * j + 1];
        r.a[6 + j] = c.z * Lm.a[3 * j + 2];
    }
    return r;
}
DEV V3 cmul(V3 c, V3 v) { return v3(c.x * v.x, c.y * v.y, c.z * v.z); }

// A.T() & B
DEV M3 mulTM(const M3& A, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = A.a[i] * B.a[j] + A.a[3 + i] * B.a[3 + j] + A.a[6 + i] * B.a[6 + j];
    return r;
}
// A & B.T()
DEV M3 mulBT(const M3& A, const M3& B)
{
    M3 r;
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
        for (int j = 0; j < 3; j++)
            r.a[3 * i + j] = A.a[3 * i] * B.a[3 * j] + A.a[3 * i + 1] * B.a[3 * j + 1] + A.a[3 * i + 2] * B.a[3 * j + 2];
    return r;
}

// Symmetric 3x3 tensor (CQW = Lam C Lam^T and CMTheta are symmetric by construction): 6 entries.
struct Sym3 {
    double xx, xy, xz, yy, yz, zz;
};
DEV Sym3 operator*(double s, const Sym3& A)
{
    Sym3 r;
    r.xx = s * A.xx; r.xy = s * A.xy; r.xz = s * A.xz; r.yy = s * A.yy; r.yz = s * A.yz; r.zz = s * A.zz;
    return r;
}
// Lam & (diag(c) & Lam.T()), upper triangle with the operation order of conjDiag
DEV Sym3 conjDiagSym(const M3& Lm, V3 c)
{
    const double a0 = c.x * Lm.a[0], a1 = c.y * Lm.a[1], a2 = c.z * Lm.a[2];
    const double b0 = c.x * Lm.a[3], b1 = c.y * Lm.a[4], b2 = c.z * Lm.a[5];
    const double c0 = c.x * Lm.a[6], c1 = c.y * Lm.a[7], c2 = c.z * Lm.a[8];
    Sym3 r;
    r.xx = Lm.a[0] * a0 + Lm.a[1] * a1 + Lm.a[2] * a2;
    r.xy = Lm.a[0] * b0 + Lm.a[1] * b1 + Lm.a[2] * b2;
    r.xz = Lm.a[0] * c0 + Lm.a[1] * c1 + Lm.a[2] * c2;
    r.yy = Lm.a[3] * b0 + Lm.a[4] * b1 + Lm.a[5] * b2;
    r.yz = Lm.a[3] * c0 + Lm.a[4] * c1 + Lm.a[5] * c2;
    r.zz = Lm.a[6] * c0 + Lm.a[7] * c1 + Lm.a[8] * c2;
    return r;
}
DEV M3 full(const Sym3& A)
{
    M3 r;
    r.a[0] = A.xx; r.a[1] = A.xy; r.a[2] = A.xz;
    r.a[3] = A.xy; r.a[4] = A.yy; r.a[5] = A.yz;
    r.a[6] = A.xz; r.a[7] = A.yz; r.a[8] = A.zz;
    return r;
}
DEV V3 symMul(const Sym3& A, V3 v)
{
    return v3(A.xx * v.x + A.xy * v.y + A.xz * v.z, A.xy * v.x + A.yy * v.y + A.yz * v.z,
              A.xz * v.x + A.yz * v.y + A.zz * v.z);
}
// A & spinTensor(t) for symmetric A
DEV M3 mulSpinR(const Sym3& A, V3 t) { return mulSpinR(full(A), t); }

// 1/x to <= 1 ulp without the IEEE-division slow path: MUFU.RCP64H seed + two Newton steps.  The compiler's
// own division guards denormal / huge operands with a CALL to a slow-path subroutine; a call site inside
// the sweep loops pins registers and forces spills around it.  The operands here (determinants of the
// diagonal blocks, |theta| + SMALL) are ordinary normal numbers; 0 -> inf, NaN -> NaN still propagate.
DEV double rcpFast(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(fma(-x, r, 1.0), r, r);
    r = fma(fma(-x, r, 1.0), r, r);
    return r;
}

// OpenFOAM inv(tensor): adjugate / determinant
DEV M3 inv(const M3& A)
{
    const double xx = A.a[0], xy = A.a[1], xz = A.a[2], yx = A.a[3], yy = A.a[4], yz = A.a[5], zx = A.a[6],
                 zy = A.a[7], zz = A.a[8];
    const double det = xx * yy * zz + xy * yz * zx + xz * yx * zy - xx * yz * zy - xy * yx * zz - xz * yy * zx;
    const double id = rcpFast(det);
    M3 r;
    r.a[0] = (yy * zz - zy * yz) * id; r.a[1] = (xz * zy - xy * zz) * id; r.a[2] = (xy * yz - xz * yy) * id;
    r.a[3] = (zx * yz - yx * zz) * id; r.a[4] = (xx * zz - xz * zx) * id; r.a[5] = (yx * xz - xx * yz) * id;
    r.a[6] = (yx * zy - yy * zx) * id; r.a[7] = (xy * zx - xx * zy) * id; r.a[8] = (xx * yy - yx * xy) * id;
    return r;
}

// inv(tensor) that also returns the determinant and the squared Frobenius norm, for the pivot guard of the block eliminations
DEV M3 invDet(const M3& A, double& det, double& frob2)
{
    const double xx = A.a[0], xy = A.a[1], xz = A.a[2], yx = A.a[3], yy = A.a[4], yz = A.a[5], zx = A.a[6],
                 zy = A.a[7], zz = A.a[8];
    det = xx * yy * zz + xy * yz * zx + xz * yx * zy - xx * yz * zy - xy * yx * zz - xz * yy * zx;
    double f = xx * xx;
#pragma unroll
    for (int k = 1; k < 9; k++) f = fma(A.a[k], A.a[k], f);
    frob2 = f;
    const double id = rcpFast(det);
    M3 r;
    r.a[0] = (yy * zz - zy * yz) * id; r.a[1] = (xz * zy - xy * zz) * id; r.a[2] = (xy * yz - xz * yy) * id;
    r.a[3] = (zx * yz - yx * zz) * id; r.a[4] = (xx * zz - xz * zx) * id; r.a[5] = (yx * xz - xx * yz) * id;
    r.a[6] = (yx * zy - yy * zx) * id; r.a[7] = (xy * zx - xx * zy) * id; r.a[8] = (xx * yy - yx * xy) * id;
    return r;
}
// A 3x3 diagonal sub-block whose determinant is small against the cube of its size is ill-conditioned
// (cond ~ |A|^3 / |det|): the adjugate inverse then loses digits (or divides by zero) where a pivoted LU of the whole
// 6x6 block -- what the reference's SparseLU does (EigenOF.C:292-297) -- still solves.  Size = the Frobenius norm
// (9 FMAs; an FP64 max costs a compare and two selects per entry): det^2 <= guard^2 (|A|_F^2 / 3)^3, which for a
// diagonal block diag(a, a, a eps) is eps <= guard.  NaN counts as "small".
#define BEAM_PIVOT_GUARD 1e-3
DEV bool smallDet(double det, double frob2)
{
    const double s = frob2 * (1.0 / 3.0);
    return !(det * det > (BEAM_PIVOT_GUARD * BEAM_PIVOT_GUARD) * (s * s * s));
}

// Rodrigues, exactly the reference's plain formula (phi = |theta| + SMALL, no series branch):
// R = I + (sin phi / phi) S + ((1 - cos phi)/phi^2) S&S      pseudoVector.C:250-264
DEV M3 rotationMatrix(V3 th)
{
    const double phi = mag(th) + BF_SMALL;
    double s, c;
    sincos(phi, &s, &c);
    const double c1 = s / phi, c2 = (1.0 - c) / (phi * phi);
    M3 R;
    // S&S = th th^T - |th|^2 I
    const double xx = th.x * th.x, yy = th.y * th.y, zz = th.z * th.z;
    R.a[0] = 1.0 + c2 * (-(yy + zz));
    R.a[4] = 1.0 + c2 * (-(xx + zz));
    R.a[8] = 1.0 + c2 * (-(xx + yy));
    const double xy = c2 * (th.x * th.y), xz = c2 * (th.x * th.z), yz = c2 * (th.y * th.z);
    R.a[1] = xy - c1 * th.z; R.a[3] = xy + c1 * th.z;
    R.a[2] = xz + c1 * th.y; R.a[6] = xz - c1 * th.y;
    R.a[5] = yz - c1 * th.x; R.a[7] = yz + c1 * th.x;
    return R;
}

// Tangent operator DT^T applied to g (Evolve.C:806-829):
// DT = (sin/phi) I + ((1 - sin/phi)/phi^2) th th^T + ((1-cos)/phi^2) S(th);  returns DT.T() & g
DEV V3 tangentT(V3 th, V3 g)
{
    const double phi = mag(th) + BF_SMALL;
    double s, c;
    sincos(phi, &s, &c);
    const double sp = s / phi;
    const double c2 = (1.0 - sp) / (phi * phi), c3 = (1.0 - c) / (phi * phi);
    // DT^T g = sp g + c2 th (th.g) - c3 (th x g)
    const double tg = dot(th, g);
    const V3 x = cross(th, g);
    return v3(sp * g.x + c2 * th.x * tg - c3 * x.x, sp * g.y + c2 * th.y * tg - c3 * x.y,
              sp * g.z + c2 * th.z * tg - c3 * x.z);
}

// ---------------------------------------------------------------------------
// Unit quaternions: the device representation of Lambda (cells) and Lambdaf (faces).
// The reference keeps the 9 entries of each rotation tensor and multiplies increments in
// (Lambda = R(DTheta) & Lambda, Evolve.C:833-841); a rotation has 3 degrees of freedom, so the
// device keeps (w, x, y, z) -- 4 doubles instead of 9 per tensor per read and per write -- and
// rebuilds the tensor in registers.  R(q1 q0) = R(q1) R(q0) exactly; the stored state differs from
// the reference's running matrix product by rounding only (both drift by O(eps) per update; the
// quaternion is renormalised each time, the matrix product is not).
// ---------------------------------------------------------------------------
struct Q4 {
    double w, x, y, z;
};
DEV Q4 q4(double w, double x, double y, double z) { Q4 q; q.w = w; q.x = x; q.y = y; q.z = z; return q; }
DEV M3 quatToMat(Q4 q)
{
    const double x2 = q.x + q.x, y2 = q.y + q.y, z2 = q.z + q.z;
    const double xx = q.x * x2, yy = q.y * y2, zz = q.z * z2;
    const double xy = q.x * y2, xz = q.x * z2, yz = q.y * z2;
    const double wx = q.w * x2, wy = q.w * y2, wz = q.w * z2;
    M3 R;
    R.a[0] = 1.0 - (yy + zz); R.a[1] = xy - wz;         R.a[2] = xz + wy;
    R.a[3] = xy + wz;         R.a[4] = 1.0 - (xx + zz); R.a[5] = yz - wx;
    R.a[6] = xz - wy;         R.a[7] = yz + wx;         R.a[8] = 1.0 - (xx + yy);
    return R;
}
// R(quatMul(a, b)) = R(a) & R(b)
DEV Q4 quatMul(Q4 a, Q4 b)
{
    return q4(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x, a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w);
}
DEV Q4 quatNormalize(Q4 q)
{
    const double n = 1.0 / sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return q4(q.w * n, q.x * n, q.y * n, q.z * n);
}
// Half-angle form of the reference's Rodrigues formula (pseudoVector.C:250-264, phi = |theta| + SMALL):
// with s = sin(phi/2), c = cos(phi/2):  sin(phi)/phi = 2 s c / phi,  (1 - cos phi)/phi^2 = 2 (s/phi)^2,
// so q = (c, (s/phi) theta) gives R(q) = I + (sin phi/phi) S + ((1 - cos phi)/phi^2) S&S.
struct HalfAngle {
    double phi, s, c;
};
DEV HalfAngle halfAngle(V3 th)
{
    HalfAngle h;
    h.phi = mag(th) + BF_SMALL;
    sincos(0.5 * h.phi, &h.s, &h.c);
    return h;
}
DEV Q4 quatFromRotVec(V3 th, const HalfAngle& h)
{
    const double k = h.s / h.phi;
    return q4(h.c, k * th.x, k * th.y, k * th.z);
}
// Tangent operator DT^T applied to g with the trigonometry taken from the half angle.
DEV V3 tangentT(V3 th, V3 g, const HalfAngle& h)
{
    const double ip = 1.0 / h.phi;
    const double sp = 2.0 * h.s * h.c * ip;           // sin(phi)/phi
    const double c3 = 2.0 * (h.s * ip) * (h.s * ip);  // (1 - cos phi)/phi^2
    const double c2 = (1.0 - sp) * (ip * ip);
    const double tg = dot(th, g);
    const V3 x = cross(th, g);
    return v3(sp * g.x + c2 * th.x * tg - c3 * x.x, sp * g.y + c2 * th.y * tg - c3 * x.y,
              sp * g.z + c2 * th.z * tg - c3 * x.z);
}
// Rotation tensor -> unit quaternion (Shepperd's branch on the largest diagonal combination);
// used only when the host sets Lambda / Lambdaf.
DEV Q4 matToQuat(const M3& R)
{
    const double tr = R.a[0] + R.a[4] + R.a[8];
    Q4 q;
    if (tr > 0.0) {
        const double s = sqrt(tr + 1.0) * 2.0;
        q = q4(0.25 * s, (R.a[7] - R.a[5]) / s, (R.a[2] - R.a[6]) / s, (R.a[3] - R.a[1]) / s);
    } else if (R.a[0] > R.a[4] && R.a[0] > R.a[8]) {
        const double s = sqrt(1.0 + R.a[0] - R.a[4] - R.a[8]) * 2.0;
        q = q4((R.a[7] - R.a[5]) / s, 0.25 * s, (R.a[1] + R.a[3]) / s, (R.a[2] + R.a[6]) / s);
    } else if (R.a[4] > R.a[8]) {
        const double s = sqrt(1.0 + R.a[4] - R.a[0] - R.a[8]) * 2.0;
        q = q4((R.a[2] - R.a[6]) / s, (R.a[1] + R.a[3]) / s, 0.25 * s, (R.a[5] + R.a[7]) / s);
    } else {
        const double s = sqrt(1.0 + R.a[8] - R.a[0] - R.a[4]) * 2.0;
        q = q4((R.a[3] - R.a[1]) / s, (R.a[2] + R.a[6]) / s, (R.a[5] + R.a[7]) / s, 0.25 * s);
    }
    return quatNormalize(q);
}

// ---------------------------------------------------------------------------
// Dense 6x6 solve with partial pivoting inside the block, NR right-hand sides,
// everything in registers: the pivot row is brought up by predicated swaps so no
// array is ever indexed dynamically.  On return B holds A^{-1} B.
// ---------------------------------------------------------------------------
template <int NR>
DEV void solve6(double (&A)[6][6], double (&B)[6][NR])
{
    double piv[6];
#pragma unroll
    for (int k = 0; k < 6; k++) {
        int p = k;
        double mx = fabs(A[k][k]);
#pragma unroll
        for (int r = k + 1; r < 6; r++) {
            const double v = fabs(A[r][k]);
            if (v > mx) { mx = v; p = r; }
        }
#pragma unroll
        for (int r = k + 1; r < 6; r++) {
            const bool sw = (p == r);
#pragma unroll
            for (int c = k; c < 6; c++) {
                const double t0 = A[k][c], t1 = A[r][c];
                A[k][c] = sw ? t1 : t0;
                A[r][c] = sw ? t0 : t1;
            }
#pragma unroll
            for (int c = 0; c < NR; c++) {
                const double t0 = B[k][c], t1 = B[r][c];
                B[k][c] = sw ? t1 : t0;
                B[r][c] = sw ? t0 : t1;
            }
        }
        const double ip = 1.0 / A[k][k];
        piv[k] = ip;
#pragma unroll
        for (int r = k + 1; r < 6; r++) {
            const double m = A[r][k] * ip;
#pragma unroll
            for (int c = k + 1; c < 6; c++) A[r][c] -= m * A[k][c];
#pragma unroll
            for (int c = 0; c < NR; c++) B[r][c] -= m * B[k][c];
        }
    }
#pragma unroll
    for (int k = 5; k >= 0; k--) {
#pragma unroll
        for (int c = 0; c < NR; c++) {
            double s = B[k][c];
#pragma unroll
            for (int j = k + 1; j < 6; j++) s -= A[k][j] * B[j][c];
            B[k][c] = s * piv[k];
        }
    }
}

// The same solve for code that must stay in registers: nvcc recognises "find the pivot row p, then exchange rows k and p"
// and indexes the arrays with the run-time p, which sends them to local memory (fine in the out-of-line fallbacks that use
// solve6, ruinous in a kernel whose whole work is this solve).  Here the largest pivot is bubbled up instead: row k is
// compared with each row below it in turn and exchanged when that one has the larger entry, so every select depends on
// values only.  The permutation differs from solve6's, the pivot (the largest entry of the column) does not.
// (The pivot steps are template-recursive: with "#pragma unroll" alone nvcc re-rolls the outer loop of this size and
// walks the rows through a pointer into local memory.)
template <int NR, int K>
DEV void solve6regStep(double (&A)[6][6], double (&B)[6][NR], double (&piv)[6])
{
    if constexpr (K < 6) {
#pragma unroll
        for (int r = K + 1; r < 6; r++) {
            const bool sw = fabs(A[r][K]) > fabs(A[K][K]);
#pragma unroll
            for (int c = K; c < 6; c++) {
                const double t0 = A[K][c], t1 = A[r][c];
                A[K][c] = sw ? t1 : t0;
                A[r][c] = sw ? t0 : t1;
            }
#pragma unroll
            for (int c = 0; c < NR; c++) {
                const double t0 = B[K][c], t1 = B[r][c];
                B[K][c] = sw ? t1 : t0;
                B[r][c] = sw ? t0 : t1;
            }
        }
        const double ip = rcpFast(A[K][K]);
        piv[K] = ip;
#pragma unroll
        for (int r = K + 1; r < 6; r++) {
            const double m = A[r][K] * ip;
#pragma unroll
            for (int c = K + 1; c < 6; c++) A[r][c] -= m * A[K][c];
#pragma unroll
            for (int c = 0; c < NR; c++) B[r][c] -= m * B[K][c];
        }
        solve6regStep<NR, K + 1>(A, B, piv);
    }
}
template <int NR, int K>
DEV void solve6regBack(double (&A)[6][6], double (&B)[6][NR], const double (&piv)[6])
{
    if constexpr (K >= 0) {
#pragma unroll
        for (int c = 0; c < NR; c++) {
            double s = B[K][c];
#pragma unroll
            for (int j = K + 1; j < 6; j++) s -= A[K][j] * B[j][c];
            B[K][c] = s * piv[K];
        }
        solve6regBack<NR, K - 1>(A, B, piv);
    }
}
template <int NR>
DEV void solve6reg(double (&A)[6][6], double (&B)[6][NR])
{
    double piv[6];
    solve6regStep<NR, 0>(A, B, piv);
    solve6regBack<NR, 5>(A, B, piv);
}

// ---------------------------------------------------------------------------
// 6x6 solve by 2x2 block elimination over 3x3 blocks, D = [[A, B], [C, E]]:
//   X2 = S^{-1} (R2 - C A^{-1} R1),  X1 = A^{-1} R1 - A^{-1} B X2,  S = E - C A^{-1} B,
// with the 3x3 inverses by adjugate/determinant (the formula OpenFOAM's inv(tensor) uses).
// No row exchanges, hence no select instructions and no dynamic register indexing.  The
// ordering is "pivoted by physics": A is the translational block (-2 CQW/dZ minus the mass
// term, always definite) and S the condensed rotational block.  On return R = D^{-1} R.
// ---------------------------------------------------------------------------
template <int NR>
DEV void solve6blk(const double (&D)[6][6], double (&R)[6][NR])
{
    M3 A, Bm, Cm, E;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++) {
            A.a[3 * r + c] = D[r][c];
            Bm.a[3 * r + c] = D[r][3 + c];
            Cm.a[3 * r + c] = D[3 + r][c];
            E.a[3 * r + c] = D[3 + r][3 + c];
        }
    const M3 Ai = inv(A);
    const M3 AiB = mul(Ai, Bm);
    const M3 Si = inv(E - mul(Cm, AiB));
#pragma unroll
    for (int c = 0; c < NR; c++) {
        const V3 y1 = mul(Ai, v3(R[0][c], R[1][c], R[2][c]));
        const V3 z2 = v3(R[3][c], R[4][c], R[5][c]) - mul(Cm, y1);
        const V3 x2 = mul(Si, z2);
        const V3 x1 = y1 - mul(AiB, x2);
        R[0][c] = x1.x; R[1][c] = x1.y; R[2][c] = x1.z;
        R[3][c] = x2.x; R[4][c] = x2.y; R[5][c] = x2.z;
    }
}
