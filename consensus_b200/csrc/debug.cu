#include "engine.h"
#include "kernels.cuh"
using namespace sbv;
// debug.cu — arithmetic-layer test hooks (used only by tests/; not part of include/sbv.h).
// Operands are little-endian 32-bit limb arrays, 2N limbs per slot (unused limbs zero).
namespace {
template <class C>
__global__ void k_debug_op(int op, uint32_t n, const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, uint32_t *__restrict__ out) {
    constexpr int N = C::N;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t x[N], y[N], u[N], v[N], r0[N], r1[N];
    for (int k = 0; k < N; k++) { x[k] = a[i * 2 * N + k]; y[k] = a[i * 2 * N + N + k]; u[k] = b[i * 2 * N + k]; v[k] = b[i * 2 * N + N + k]; r0[k] = 0; r1[k] = 0; }
    uint32_t rr[N], one[N], plain1[N];
    C::get_rr_p(rr); C::get_one(one);
    for (int k = 0; k < N; k++) plain1[k] = (k == 0);
    if (op == 0) C::fmul(r0, x, u);
    else if (op == 1) C::fadd(r0, x, u);
    else if (op == 2) C::fsub(r0, x, u);
    else if (op == 3) C::nmul(r0, x, u);
    else if (op == 4) f_inv<C>(r0, x);
    else if (op == 8) n_inv<C>(r0, x);
    else if (op == 9) C::fsqr(r0, x);
    else if (op >= 5 && op <= 7) {
        // affine plain (x,y) [+ (u,v)] -> Montgomery Jacobian -> op -> affine plain
        Jac<C> P;
        C::fmul(P.X, x, rr); C::fmul(P.Y, y, rr); mp_copy<N>(P.Z, one);
        uint32_t um[N], vm[N];
        C::fmul(um, u, rr); C::fmul(vm, v, rr);
        if (op == 5) pt_double<C>(P);
        else if (op == 6) {
            // general add with a non-trivial Z2: scale (u,v) by z=3 -> (9u, 27v, 3)
            uint32_t z[N], z2[N], z3[N], t[N];
            C::fadd(t, one, one); C::fadd(z, t, one);
            C::fsqr(z2, z); C::fmul(z3, z2, z);
            C::fmul(um, um, z2); C::fmul(vm, vm, z3);
            pt_double<C>(P);  // make Z1 non-trivial as well: P = 2*(x,y)
            pt_add<C, false>(P, um, vm, z, false, false);
        } else pt_add<C, true>(P, um, vm, one, false, false);
        if (mp_is_zero<N>(P.Z)) { for (int k = 0; k < N; k++) { r0[k] = 0; r1[k] = 0; } }
        else {
            uint32_t zi[N], zi2[N], zi3[N];
            f_inv<C>(zi, P.Z);
            C::fsqr(zi2, zi); C::fmul(zi3, zi2, zi);
            C::fmul(r0, P.X, zi2); C::fmul(r1, P.Y, zi3);
            C::fmul(r0, r0, plain1); C::fmul(r1, r1, plain1);  // out of Montgomery form
        }
    }
    for (int k = 0; k < N; k++) { out[i * 2 * N + k] = r0[k]; out[i * 2 * N + N + k] = r1[k]; }
}
}  // namespace

extern "C" int sbv_debug_op(sbv_engine *e, uint8_t curve, int op, size_t n, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    if (!e || curve > SBV_P384) return SBV_ERR_ARG;
    std::lock_guard<std::mutex> lk(e->mu);
    Dev &d = e->devs[0];
    CU(e, cudaSetDevice(d.ordinal));
    const size_t N = curve == SBV_P256 ? 8 : 12, bytes = n * 2 * N * 4;
    int rc = sbv_ensure_scratch(e, d, 3 * bytes + 1024);
    if (rc) return rc;
    uint32_t *da = (uint32_t *)d.d_scratch, *db = da + n * 2 * N, *dout = db + n * 2 * N;
    CU(e, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, d.stream));
    CU(e, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, d.stream));
    if (curve == SBV_P256) k_debug_op<P256><<<(uint32_t)((n + 63) / 64), 64, 0, d.stream>>>(op, (uint32_t)n, da, db, dout);
    else k_debug_op<P384><<<(uint32_t)((n + 63) / 64), 64, 0, d.stream>>>(op, (uint32_t)n, da, db, dout);
    CU(e, cudaGetLastError());
    CU(e, cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, d.stream));
    CU(e, cudaStreamSynchronize(d.stream));
    return SBV_OK;
}
