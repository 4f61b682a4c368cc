// sinf_glibc.h -- restatement of glibc 2.39's sinf (sysdeps/ieee754/flt-32/s_sinf.c, sincosf.h,
// sincosf_data.c; the FMA build that x86-64 hosts with FMA select through ifunc), shared by the
// device-side filter design and a host twin the CPU test-suite holds against libm's sinf.
// The reference computes its sinc with f32::sin, i.e. the platform's sinf (src/window.rs:29-36), so
// the table's bits depend on it.  Everything is f64 arithmetic with one final rounding to f32.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define RSB_SINF_HD __host__ __device__
#define RSB_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RSB_MUL(a, b) __dmul_rn((a), (b))
#define RSB_D2F(a) __double2float_rn(a)
#define RSB_D2I_RZ(a) __double2int_rz(a)
#define RSB_LL2D(a) __ll2double_rn(a)
#else
#if defined(__CUDACC__)
#define RSB_SINF_HD __host__ __device__
#else
#define RSB_SINF_HD
#endif
#define RSB_FMA(a, b, c) std::fma((a), (b), (c))
#define RSB_MUL(a, b) ((a) * (b))
#define RSB_D2F(a) static_cast<float>(a)
#define RSB_D2I_RZ(a) static_cast<int32_t>(a)
#define RSB_LL2D(a) static_cast<double>(a)
#endif

namespace rsb {

struct SinCosTab {
    double sign[4];
    double hpi_inv, hpi, c0, c1, c2, c3, c4, s1, s2, s3;
};

RSB_SINF_HD inline const SinCosTab &sincos_tab(int i) {
    // __sincosf_table (hpi_inv prescaled by 2^24: the non-TOINT_INTRINSICS build)
    static const
#if defined(__CUDA_ARCH__)
        __device__
#endif
        SinCosTab tab[2] = {
            {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, 0x1p0, -0x1.ffffffd0c621cp-2,
             0x1.55553e1068f19p-5, -0x1.6c087e89a359dp-10, 0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
             0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13},
            {{1.0, -1.0, -1.0, 1.0}, 0x1.45F306DC9C883p+23, 0x1.921FB54442D18p0, -0x1p0, 0x1.ffffffd0c621cp-2,
             -0x1.55553e1068f19p-5, 0x1.6c087e89a359dp-10, -0x1.99343027bf8c3p-16, -0x1.555545995a603p-3,
             0x1.1107605230bc4p-7, -0x1.994eb3774cf24p-13}};
    return tab[i];
}

RSB_SINF_HD inline uint32_t inv_pio4(uint32_t i) {
    // __inv_pio4: 2/pi in 32-bit pieces, overlapping by 24 bits
    static const
#if defined(__CUDA_ARCH__)
        __device__
#endif
        uint32_t t[24] = {0xa2,       0xa2f9,     0xa2f983,   0xa2f9836e, 0xf9836e4e, 0x836e4e44, 0x6e4e4415, 0x4e441529,
                          0x441529fc, 0x1529fc27, 0x29fc2757, 0xfc2757d1, 0x2757d1f5, 0x57d1f534, 0xd1f534dd, 0xf534ddc0,
                          0x34ddc0db, 0xddc0db62, 0xc0db6295, 0xdb629599, 0x6295993c, 0x95993c43, 0x993c4390, 0x3c439041};
    return t[i];
}

RSB_SINF_HD inline uint32_t f32_bits(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    return u;
}
RSB_SINF_HD inline uint32_t abstop12(float x) { return (f32_bits(x) >> 20) & 0x7ffu; }

RSB_SINF_HD inline float sinf_poly(double x, double x2, const SinCosTab &p, int n) {
    if ((n & 1) == 0) {
        const double x3 = RSB_MUL(x, x2);
        const double s1 = RSB_FMA(x2, p.s3, p.s2);
        const double x7 = RSB_MUL(x3, x2);
        const double s = RSB_FMA(x3, p.s1, x);
        return RSB_D2F(RSB_FMA(x7, s1, s));
    }
    const double x4 = RSB_MUL(x2, x2);
    const double c2 = RSB_FMA(x2, p.c4, p.c3);
    const double c1 = RSB_FMA(x2, p.c1, p.c0);
    const double x6 = RSB_MUL(x4, x2);
    const double c = RSB_FMA(x4, p.c2, c1);
    return RSB_D2F(RSB_FMA(x6, c2, c));
}

// finite arguments only (the table's stay below 2^8)
RSB_SINF_HD inline float sinf_glibc(float y) {
    double x = static_cast<double>(y);
    const SinCosTab &p0 = sincos_tab(0);
    if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {            // |y| < pi/4
        if (abstop12(y) < abstop12(0x1p-12f)) return y;
        return sinf_poly(x, RSB_MUL(x, x), p0, 0);
    }
    if (abstop12(y) < abstop12(120.0f)) {                     // reduce_fast
        const double r = RSB_MUL(x, p0.hpi_inv);
        const int n = (RSB_D2I_RZ(r) + 0x800000) >> 24;
        x = RSB_FMA(-static_cast<double>(n), p0.hpi, x);
        const double s = p0.sign[n & 3];
        return sinf_poly(RSB_MUL(x, s), RSB_MUL(x, x), sincos_tab((n & 2) ? 1 : 0), n);
    }
    // reduce_large
    uint32_t xi = f32_bits(y);
    const int sign = static_cast<int>(xi >> 31);
    const uint32_t a0 = (xi >> 26) & 15u;
    const int shift = static_cast<int>((xi >> 23) & 7u);
    xi = (xi & 0xffffffu) | 0x800000u;
    xi <<= shift;
    uint64_t res0 = static_cast<uint64_t>(static_cast<uint32_t>(xi * inv_pio4(a0)));
    const uint64_t res1 = static_cast<uint64_t>(xi) * inv_pio4(a0 + 4);
    const uint64_t res2 = static_cast<uint64_t>(xi) * inv_pio4(a0 + 8);
    res0 = (res2 >> 32) | (res0 << 32);
    res0 += res1;
    const uint64_t nn = (res0 + (1ull << 61)) >> 62;
    res0 -= nn << 62;
    const int n = static_cast<int>(nn);
    x = RSB_MUL(RSB_LL2D(static_cast<long long>(res0)), 0x1.921FB54442D18p-62);
    const double s = p0.sign[(n + sign) & 3];
    return sinf_poly(RSB_MUL(x, s), RSB_MUL(x, x), sincos_tab(((n + sign) & 2) ? 1 : 0), n);
}

}  // namespace rsb
