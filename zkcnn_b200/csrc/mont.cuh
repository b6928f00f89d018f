// Montgomery-form prime-field arithmetic for BLS12-381, usable from host and device code.
//
// Replaces (value-identically) the arithmetic that the reference delegates to herumi/mcl v1.22:
//   Fr = 4x64-bit Montgomery limbs, R = 2^256  (mcl_fp_mont4L, mcl/src/asm/x86-64.s:1687)
//   Fp = 6x64-bit Montgomery limbs, R = 2^384  (mcl_fp_mont6L, mcl/src/asm/x86-64.s:3524)
// The in-memory form is bit-identical to mcl's (little-endian limbs, fully reduced), so a `std::vector<Fr>` of the
// reference can be handed to the C ABI without conversion.  Here the limbs are 32-bit because
// the B200 integer pipe is 32 bits wide.
//
// Two implementations of the multiplier exist: a portable one (plain C, also the host path) and an inline-PTX one
// using mad.lo.cc/madc.hi.cc carry chains.  Both are CIOS with a (N+1)-limb running accumulator; because both moduli
// leave >= 1 spare top bit, the accumulator never needs limb N+1 ("field arithmetic").
#pragma once
#include "zk_platform.cuh"
#include "bls12_381_constants.cuh"

namespace zk {

#if ZK_ON_DEVICE && !defined(ZK_PORTABLE_FIELD)
#define ZK_FIELD_PTX 1
#else
#define ZK_FIELD_PTX 0
#endif

#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
namespace ptx {
__device__ __forceinline__ uint32_t add_cc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t addc_cc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t addc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t sub_cc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t subc_cc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t subc(uint32_t x, uint32_t y) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t mul_lo(uint32_t x, uint32_t y) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t mul_hi(uint32_t x, uint32_t y) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__device__ __forceinline__ uint32_t mad_lo_cc(uint32_t x, uint32_t y, uint32_t z) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
__device__ __forceinline__ uint32_t madc_lo_cc(uint32_t x, uint32_t y, uint32_t z) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
__device__ __forceinline__ uint32_t mad_hi_cc(uint32_t x, uint32_t y, uint32_t z) { uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
__device__ __forceinline__ uint32_t madc_hi_cc(uint32_t x, uint32_t y, uint32_t z) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
__device__ __forceinline__ uint32_t madc_hi(uint32_t x, uint32_t y, uint32_t z) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(z)); return r; }
}  // namespace ptx
#endif

#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
// run-time copies of the moduli (not `const`: the compiler must not turn the limbs into immediates, see mul_wide)
static __device__ __constant__ uint32_t d_fr_MOD_rt[9] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u, 0xffffffffu};
static __device__ __constant__ uint32_t d_fp_MOD_rt[13] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u, 0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau, 0xfffcfffdu};
#endif
struct fr_cfg {
    static constexpr int N = 8;
    // r = ... ffffffff 00000001 in its two low limbs: m * r[0] = m and m * r[1] = (m << 32) - m need no multiplier
    // (wide_reduce_cmad).  Measured on B200 it does not pay: 16 wide multiply-adds fewer but 48 integer-ALU instructions
    // more per multiplication leave the throughput at 65-67 Gmul/s either way and the fold kernel 2 % slower
    // (tools/mulbench.cu, DESIGN.md "cost model"), so the generic path stays the default.
    static constexpr bool LOW_LIMBS_SPECIAL = false;
    static constexpr uint32_t INV = fr_params::INV;
    static ZK_HD __forceinline__ const uint32_t *mod() { return ZK_C(fr_MOD); }
#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
    static __device__ __forceinline__ const uint32_t *mod_rt() { return d_fr_MOD_rt; }
#endif
    static ZK_HD __forceinline__ const uint32_t *one() { return ZK_C(fr_ONE); }
    static ZK_HD __forceinline__ const uint32_t *r2() { return ZK_C(fr_R2); }
    static ZK_HD __forceinline__ const uint32_t *r3() { return ZK_C(fr_R3); }
    static ZK_HD __forceinline__ const uint32_t *modm2() { return ZK_C(fr_MODM2); }
    static ZK_HD __forceinline__ const uint32_t *half() { return ZK_C(fr_HALF); }
};
struct fp_cfg {
    static constexpr int N = 12;
    static constexpr bool LOW_LIMBS_SPECIAL = false;
    static constexpr uint32_t INV = fp_params::INV;
    static ZK_HD __forceinline__ const uint32_t *mod() { return ZK_C(fp_MOD); }
#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
    static __device__ __forceinline__ const uint32_t *mod_rt() { return d_fp_MOD_rt; }
#endif
    static ZK_HD __forceinline__ const uint32_t *one() { return ZK_C(fp_ONE); }
    static ZK_HD __forceinline__ const uint32_t *r2() { return ZK_C(fp_R2); }
    static ZK_HD __forceinline__ const uint32_t *r3() { return ZK_C(fp_R3); }
    static ZK_HD __forceinline__ const uint32_t *modm2() { return ZK_C(fp_MODM2); }
    static ZK_HD __forceinline__ const uint32_t *half() { return ZK_C(fp_HALF); }
};

template <class C> struct alignas(16) mont_t {
    static constexpr int N = C::N;
    uint32_t v[N];

    // ---- constructors / constants -------------------------------------------------------------------------------
    static ZK_HD __forceinline__ mont_t zero() {
        mont_t r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = 0;
        return r;
    }
    static ZK_HD __forceinline__ mont_t one() {
        mont_t r;
#pragma unroll
        for (int i = 0; i < N; ++i) r.v[i] = C::one()[i];
        return r;
    }
    // canonical integer (little-endian 32-bit limbs, < modulus) -> Montgomery form
    static ZK_HD inline mont_t from_canonical(const uint32_t *x) {
        mont_t a, r2;
#pragma unroll
        for (int i = 0; i < N; ++i) { a.v[i] = x[i]; r2.v[i] = C::r2()[i]; }
        return a * r2;
    }
    static ZK_HD inline mont_t from_u64(uint64_t x) {
        uint32_t c[N];
        for (int i = 0; i < N; ++i) c[i] = 0;
        c[0] = (uint32_t) x;
        c[1] = (uint32_t) (x >> 32);
        return from_canonical(c);
    }
    static ZK_HD inline mont_t from_i64(int64_t x) { return x < 0 ? -from_u64((uint64_t) (-x)) : from_u64((uint64_t) x); }
    // Montgomery form -> canonical integer
    ZK_HD inline void to_canonical(uint32_t *out) const {
        mont_t u;
#pragma unroll
        for (int i = 0; i < N; ++i) u.v[i] = i == 0 ? 1u : 0u;
        mont_t r = *this * u;
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = r.v[i];
    }

    // ---- predicates -----------------------------------------------------------------------------------------------
    ZK_HD __forceinline__ bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i];
        return o == 0;
    }
    ZK_HD __forceinline__ bool operator==(const mont_t &b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    ZK_HD __forceinline__ bool operator!=(const mont_t &b) const { return !(*this == b); }

    // raw little-endian comparison a >= b on limb arrays
    static ZK_HD __forceinline__ bool ge_raw(const uint32_t *a, const uint32_t *b) {
        for (int i = N - 1; i >= 0; --i) {
            if (a[i] > b[i]) return true;
            if (a[i] < b[i]) return false;
        }
        return true;
    }

    // ---- add / sub / neg --------------------------------------------------------------------------------------------
    friend ZK_HD __forceinline__ mont_t operator+(const mont_t &a, const mont_t &b) {
        mont_t s, d;
        const uint32_t *p = C::mod();
#if ZK_FIELD_PTX
        s.v[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) s.v[i] = ptx::addc_cc(a.v[i], b.v[i]);
        s.v[N - 1] = ptx::addc(a.v[N - 1], b.v[N - 1]);  // no carry out: both moduli leave a spare top bit
        d.v[0] = ptx::sub_cc(s.v[0], p[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d.v[i] = ptx::subc_cc(s.v[i], p[i]);
        uint32_t borrow = ptx::subc(0, 0);  // 0 - 0 - borrow = 0xffffffff if s < p
#pragma unroll
        for (int i = 0; i < N; ++i) s.v[i] = borrow ? s.v[i] : d.v[i];
        return s;
#else
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            c += (uint64_t) a.v[i] + b.v[i];
            s.v[i] = (uint32_t) c;
            c >>= 32;
        }
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bw += (int64_t) s.v[i] - (int64_t) p[i];
            d.v[i] = (uint32_t) bw;
            bw >>= 32;
        }
        return bw ? s : d;
#endif
    }
    friend ZK_HD __forceinline__ mont_t operator-(const mont_t &a, const mont_t &b) {
        mont_t d, s;
        const uint32_t *p = C::mod();
#if ZK_FIELD_PTX
        d.v[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d.v[i] = ptx::subc_cc(a.v[i], b.v[i]);
        uint32_t borrow = ptx::subc(0, 0);
        s.v[0] = ptx::add_cc(d.v[0], p[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) s.v[i] = ptx::addc_cc(d.v[i], p[i]);
        s.v[N - 1] = ptx::addc(d.v[N - 1], p[N - 1]);
#pragma unroll
        for (int i = 0; i < N; ++i) d.v[i] = borrow ? s.v[i] : d.v[i];
        return d;
#else
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bw += (int64_t) a.v[i] - (int64_t) b.v[i];
            d.v[i] = (uint32_t) bw;
            bw >>= 32;
        }
        if (!bw) return d;
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            c += (uint64_t) d.v[i] + p[i];
            s.v[i] = (uint32_t) c;
            c >>= 32;
        }
        return s;
#endif
    }
    ZK_HD __forceinline__ mont_t operator-() const { return is_zero() ? *this : (zero() - *this); }
    // a - b + p in (0, 2p): a difference that is NOT brought back below p (16 instructions instead of 25).  Valid as the
    // SECOND operand of a multiplication (the product still ends below 2p before the final conditional subtraction) and
    // as either operand of lazy_acc_t::mac; never stored.
    static ZK_HD __forceinline__ mont_t sub_lazy(const mont_t &a, const mont_t &b) {
        mont_t t;
        const uint32_t *p = C::mod();
#if ZK_FIELD_PTX
        t.v[0] = ptx::add_cc(a.v[0], p[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) t.v[i] = ptx::addc_cc(a.v[i], p[i]);
        t.v[N - 1] = ptx::addc(a.v[N - 1], p[N - 1]);
        t.v[0] = ptx::sub_cc(t.v[0], b.v[0]);
#pragma unroll
        for (int i = 1; i < N - 1; ++i) t.v[i] = ptx::subc_cc(t.v[i], b.v[i]);
        t.v[N - 1] = ptx::subc(t.v[N - 1], b.v[N - 1]);
#else
        uint64_t c = 0;
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            c += (uint64_t) a.v[i] + p[i];
            t.v[i] = (uint32_t) c;
            c >>= 32;
        }
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bw += (int64_t) t.v[i] - (int64_t) b.v[i];
            t.v[i] = (uint32_t) bw;
            bw >>= 32;
        }
#endif
        return t;
    }

    // ---- multiplication ---------------------------------------------------------------------------------------------
    static ZK_HD __forceinline__ void mul_portable(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        const uint32_t *p = C::mod();
        uint32_t t[N + 1];
#if !ZK_ON_DEVICE && defined(__SIZEOF_INT128__)
        // host code: the same interleaved Montgomery multiplication on 64-bit limbs (the quotient digits of a + m * p == 0 mod 2^(32N) are
        // unique, so the unreduced result equals the 32-bit loop's); ~5x faster, which the host-side finish of the opening MSMs relies on
        static_assert(N % 2 == 0, "64-bit host path packs pairs of limbs");
        {
            typedef unsigned __int128 u128;
            constexpr int M = N / 2;
            uint64_t A[M], B[M], P[M], T[M + 2];
            for (int i = 0; i < M; ++i) {
                A[i] = (uint64_t) a[2 * i] | ((uint64_t) a[2 * i + 1] << 32);
                B[i] = (uint64_t) b[2 * i] | ((uint64_t) b[2 * i + 1] << 32);
                P[i] = (uint64_t) p[2 * i] | ((uint64_t) p[2 * i + 1] << 32);
            }
            for (int i = 0; i < M + 2; ++i) T[i] = 0;
            uint64_t pinv = (uint64_t) (0u - C::INV);   // p^-1 mod 2^32 (INV is -p^-1), one Newton step doubles the precision
            pinv *= 2 - P[0] * pinv;
            const uint64_t inv64 = 0 - pinv;
            for (int i = 0; i < M; ++i) {
                u128 c = 0;
                for (int j = 0; j < M; ++j) {
                    const u128 x = (u128) A[j] * B[i] + T[j] + c;
                    T[j] = (uint64_t) x;
                    c = x >> 64;
                }
                u128 top = (u128) T[M] + c;
                T[M] = (uint64_t) top;
                T[M + 1] = (uint64_t) (top >> 64);
                const uint64_t m = T[0] * inv64;
                c = ((u128) m * P[0] + T[0]) >> 64;
                for (int j = 1; j < M; ++j) {
                    const u128 x = (u128) m * P[j] + T[j] + c;
                    T[j - 1] = (uint64_t) x;
                    c = x >> 64;
                }
                top = (u128) T[M] + c;
                T[M - 1] = (uint64_t) top;
                T[M] = T[M + 1] + (uint64_t) (top >> 64);
            }
            for (int i = 0; i < M; ++i) { t[2 * i] = (uint32_t) T[i]; t[2 * i + 1] = (uint32_t) (T[i] >> 32); }
            t[N] = (uint32_t) T[M];
        }
#else
#pragma unroll
        for (int i = 0; i <= N; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            uint64_t c = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                uint64_t x = (uint64_t) a[j] * b[i] + t[j] + c;
                t[j] = (uint32_t) x;
                c = x >> 32;
            }
            uint64_t top = (uint64_t) t[N] + c;  // fits: accumulator < 2^(32(N+1)), see header comment
            uint32_t m = t[0] * C::INV;
            c = ((uint64_t) m * p[0] + t[0]) >> 32;
#pragma unroll
            for (int j = 1; j < N; ++j) {
                uint64_t x = (uint64_t) m * p[j] + t[j] + c;
                t[j - 1] = (uint32_t) x;
                c = x >> 32;
            }
            top += c;
            t[N - 1] = (uint32_t) top;
            t[N] = (uint32_t) (top >> 32);
        }
#endif
        // result < 2p: one conditional subtraction
        uint32_t d[N];
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bw += (int64_t) t[i] - (int64_t) p[i];
            d[i] = (uint32_t) bw;
            bw >>= 32;
        }
        bool keep = bw != 0 && t[N] == 0;
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = keep ? t[i] : d[i];
    }

#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
    // t[0..N] += x[0..N-1] * y   (one low-half carry chain, one high-half carry chain)
    static __device__ __forceinline__ void mad_row(uint32_t *t, const uint32_t *x, uint32_t y) {
        t[0] = ptx::mad_lo_cc(x[0], y, t[0]);
#pragma unroll
        for (int j = 1; j < N; ++j) t[j] = ptx::madc_lo_cc(x[j], y, t[j]);
        t[N] = ptx::addc(t[N], 0);
        t[1] = ptx::mad_hi_cc(x[0], y, t[1]);
#pragma unroll
        for (int j = 1; j < N - 1; ++j) t[j + 1] = ptx::madc_hi_cc(x[j], y, t[j + 1]);
        t[N] = ptx::madc_hi(x[N - 1], y, t[N]);
    }
    static __device__ __forceinline__ void mul_ptx(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        const uint32_t *p = C::mod();
        uint32_t pm[N];
#pragma unroll
        for (int i = 0; i < N; ++i) pm[i] = p[i];
        uint32_t t[N + 2];
#pragma unroll
        for (int i = 0; i < N + 2; ++i) t[i] = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            mad_row(t, a, b[i]);
            uint32_t m = t[0] * C::INV;
            mad_row(t, pm, m);
            // divide by 2^32: t[0] is now zero
#pragma unroll
            for (int j = 0; j <= N; ++j) t[j] = t[j + 1];
        }
        uint32_t d[N];
        d[0] = ptx::sub_cc(t[0], pm[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d[i] = ptx::subc_cc(t[i], pm[i]);
        uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = borrow ? t[i] : d[i];
    }

    // ---- "wide" multiplier: even/odd column accumulators ---------------------------------------------------------------
    // The pairs  mad.lo.cc a,b,lo ; madc.hi.cc a,b,hi  below are what ptxas turns into ONE  IMAD.WIDE.U32(.X)  each, so a
    // Montgomery multiplication costs 2 N^2 wide multiply-adds plus O(N) glue instead of 4 N^2 half products and as many
    // carry fix-ups.  Two accumulators E ("even") and O ("odd", one limb higher) hold the running value  E + 2^32 O; the
    // products a[j] b_i land in E for even j and in O for odd j, so every wide result is limb-aligned with its
    // accumulator and each accumulator needs a single carry chain per row.  Dividing by 2^32 after the reduction row is
    // free: the two accumulators swap roles (the old O is the new E).
    //   acc[j], acc[j+1] += x[j] * y  for j = 0, 2, ..   (one carry chain; the carry out stays in CC.CF)
    static __device__ __forceinline__ void wide_cmad(uint32_t *acc, const uint32_t *x, uint32_t y) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(x[0]), "r"(y));
#pragma unroll
        for (int j = 2; j < N; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(x[j]), "r"(y));
    }
    //   acc[j], acc[j+1] = x[j] * y + (acc[j+2], acc[j+3])   (shift down by two limbs while accumulating; consumes CC.CF)
    static __device__ __forceinline__ void wide_madc_rshift(uint32_t *acc, const uint32_t *x, uint32_t y) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                         : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(x[j]), "r"(y), "r"(acc[j + 2]), "r"(acc[j + 3]));
        asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(acc[N - 2]), "=r"(acc[N - 1]) : "r"(x[N - 2]), "r"(y));
    }
    //   acc[j], acc[j+1] = x[j] * y
    static __device__ __forceinline__ void wide_mul(uint32_t *acc, const uint32_t *x, uint32_t y) {
#pragma unroll
        for (int j = 0; j < N; j += 2)
            asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(x[j]), "r"(y));
    }
    // one row: (E + 2^32 O) <- ((E + 2^32 O) / 2^32 [previous row's pending shift]) + a * bi + m * p, with E[0] == 0 after
    template <bool FIRST> static __device__ __forceinline__ void wide_row(uint32_t *E, uint32_t *O, const uint32_t *a, uint32_t bi,
                                                                          const uint32_t *pm) {
        if (FIRST) {
            wide_mul(O, a + 1, bi);
            wide_mul(E, a, bi);
        } else {
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(O[1]));
            wide_madc_rshift(O, a + 1, bi);
            wide_cmad(E, a, bi);
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[N - 1]));
        }
        // m = E[0] * (-p^-1) with the factor read from constant memory: for Fr it is 0xffffffff and, seen as an immediate,
        // ptxas turns the product into a negation and then no longer fuses the m * p pairs below into IMAD.WIDE
        uint32_t m;
        asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(E[0]), "r"(pm[N]));
        wide_reduce_cmad(E, O, pm, m);
    }
    // (E + 2^32 O) += m * p, which clears E[0].  No carry out of O: the running value stays below 2^(32 N + 32) (spare
    // top bit of p).  For Fr the two low limbs of the modulus are 1 and 2^32 - 1, so their products with m are formed on
    // the integer ALU (which idles while the multiplier pipe is the bottleneck): 6 instead of 8 wide multiply-adds per row.
    static __device__ __forceinline__ void wide_reduce_cmad(uint32_t *E, uint32_t *O, const uint32_t *pm, uint32_t m) {
#ifndef ZK_NO_SPECIAL_MODULUS
        if (C::LOW_LIMBS_SPECIAL) {
            const uint32_t lo1 = 0u - m, hi1 = m - (m != 0u ? 1u : 0u);   // m * (2^32 - 1) = (m << 32) - m
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(O[0]) : "r"(lo1));
            asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(O[1]) : "r"(hi1));
#pragma unroll
            for (int j = 2; j < N; j += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(O[j]), "+r"(O[j + 1]) : "r"(pm[j + 1]), "r"(m));
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(m));   // m * 1: the sum is 0 mod 2^32 by the choice of m
            asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(E[1]));
#pragma unroll
            for (int j = 2; j < N; j += 2)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(E[j]), "+r"(E[j + 1]) : "r"(pm[j]), "r"(m));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[N - 1]));
            return;
        }
#endif
        wide_cmad(O, pm + 1, m);
        wide_cmad(E, pm, m);
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(O[N - 1]));
    }
    static __device__ __forceinline__ void mul_wide(uint32_t *r, const uint32_t *a, const uint32_t *b) {
        static_assert(N % 2 == 0, "even limb count");
#ifdef ZK_WIDE_MOD_IMM
        const uint32_t *p = C::mod();
#else
        const uint32_t *p = C::mod_rt();
#endif
        uint32_t pm[N + 1];
#pragma unroll
        for (int i = 0; i < N; ++i) pm[i] = p[i];
        pm[N] = C::mod_rt()[N];   // -p^-1 mod 2^32, deliberately opaque to the compiler (see wide_row)
        uint32_t E[N], O[N];
        wide_row<true>(E, O, a, b[0], pm);
        wide_row<false>(O, E, a, b[1], pm);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            wide_row<false>(E, O, a, b[i], pm);
            wide_row<false>(O, E, a, b[i + 1], pm);
        }
        // merge: value = (E >> 32) + O
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(O[1]));
#pragma unroll
        for (int i = 1; i < N - 1; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(E[i]) : "r"(O[i + 1]));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N - 1]));
        uint32_t d[N];
        d[0] = ptx::sub_cc(E[0], pm[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d[i] = ptx::subc_cc(E[i], pm[i]);
        uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = borrow ? E[i] : d[i];
    }
    // Two independent multiplications with their rows interleaved in program order: the carry chains of the two products
    // do not depend on each other, so a single warp keeps the multiplier pipe busy while one chain waits for its carry.
    // A lone warp needs ~2340 clk per Fp multiplication against a pipe time of 1152 (tools/latbench.cu): the pair costs
    // about as much as one.  This is what the latency-bound curve kernels (few warps per SM, long dependent chains of
    // point operations) are built on.
    static __device__ __forceinline__ void mul_wide2(uint32_t *r0, const uint32_t *a0, const uint32_t *b0, uint32_t *r1, const uint32_t *a1,
                                                     const uint32_t *b1) {
        const uint32_t *p = C::mod_rt();
        uint32_t pm[N + 1];
#pragma unroll
        for (int i = 0; i <= N; ++i) pm[i] = p[i];
        uint32_t E0[N], O0[N], E1[N], O1[N];
        wide_row<true>(E0, O0, a0, b0[0], pm);
        wide_row<true>(E1, O1, a1, b1[0], pm);
        wide_row<false>(O0, E0, a0, b0[1], pm);
        wide_row<false>(O1, E1, a1, b1[1], pm);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            wide_row<false>(E0, O0, a0, b0[i], pm);
            wide_row<false>(E1, O1, a1, b1[i], pm);
            wide_row<false>(O0, E0, a0, b0[i + 1], pm);
            wide_row<false>(O1, E1, a1, b1[i + 1], pm);
        }
        wide_finish(r0, E0, O0, pm);
        wide_finish(r1, E1, O1, pm);
    }
    // merge (E >> 32) + O and subtract p once if needed
    static __device__ __forceinline__ void wide_finish(uint32_t *r, uint32_t *E, const uint32_t *O, const uint32_t *pm) {
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(E[0]) : "r"(O[1]));
#pragma unroll
        for (int i = 1; i < N - 1; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(E[i]) : "r"(O[i + 1]));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(E[N - 1]));
        uint32_t d[N];
        d[0] = ptx::sub_cc(E[0], pm[0]);
#pragma unroll
        for (int i = 1; i < N; ++i) d[i] = ptx::subc_cc(E[i], pm[i]);
        uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < N; ++i) r[i] = borrow ? E[i] : d[i];
    }

    // ---- Montgomery reduction of ONE N-limb integer: r = t / R mod p (t any value below R), fully reduced ----------------
    // The rows of mul_wide with a = 1: the a * b_i term is the single limb t[i] entering at limb 0.
    static __device__ __forceinline__ void redc_wide(uint32_t *r, const uint32_t *t) {
        const uint32_t *p = C::mod_rt();
        uint32_t pm[N + 1];
#pragma unroll
        for (int i = 0; i <= N; ++i) pm[i] = p[i];
        uint32_t E[N], O[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { E[i] = 0; O[i] = 0; }
        E[0] = t[0];
        {
            uint32_t m;
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(E[0]), "r"(pm[N]));
            wide_reduce_cmad(E, O, pm, m);
        }
#pragma unroll
        for (int i = 1; i < N; ++i) {
            uint32_t *X = (i & 1) ? O : E, *Y = (i & 1) ? E : O;   // X: aligned at limb 0 after the shift, Y: shifts down by two limbs
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(X[0]) : "r"(Y[1]));
#pragma unroll
            for (int j = 0; j < N - 2; ++j) asm volatile("addc.cc.u32 %0, %1, 0;" : "=r"(Y[j]) : "r"(Y[j + 2]));
            asm volatile("addc.u32 %0, 0, 0;" : "=r"(Y[N - 2]));
            Y[N - 1] = 0;
            asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(X[0]) : "r"(t[i]));
#pragma unroll
            for (int j = 1; j < N; ++j) asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(X[j]));
            asm volatile("addc.u32 %0, %0, 0;" : "+r"(Y[N - 1]));
            uint32_t m;
            asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(m) : "r"(X[0]), "r"(pm[N]));
            wide_reduce_cmad(X, Y, pm, m);
        }
        // after N rows the roles are: N even -> the value is (E >> 32) + O with E the accumulator whose limb 0 is zero
        if (N & 1) wide_finish(r, O, E, pm);
        else wide_finish(r, E, O, pm);
    }

    // ---- unreduced product t[0 .. 2N-1] = a * b as plain integers (the multiplication half of mul_wide) ----------------
    // Same even/odd column scheme: `t` collects the limb-aligned wide products, `odd` the ones shifted by one limb.
    //   row: odd[0..N-1] += x[1,3,..] * y with the top pair written fresh; even[0..N-1] += x[0,2,..] * y; the carry out of
    //   the even chain lands in odd[N-1], whose fresh high half has room for it.
    static __device__ __forceinline__ void wide_raw_row(uint32_t *odd, uint32_t *even, const uint32_t *x, uint32_t y) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(odd[0]), "+r"(odd[1]) : "r"(x[1]), "r"(y));
#pragma unroll
        for (int j = 2; j < N - 2; j += 2)
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(odd[j]), "+r"(odd[j + 1]) : "r"(x[j + 1]), "r"(y));
        asm volatile("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(odd[N - 2]), "=r"(odd[N - 1]) : "r"(x[N - 1]), "r"(y));
        wide_cmad(even, x, y);
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(odd[N - 1]));
    }
    static __device__ __forceinline__ void mul_raw_wide(uint32_t *t, const uint32_t *a, const uint32_t *b) {
        uint32_t odd[2 * N - 2];
        wide_mul(t, a, b[0]);
        wide_mul(odd, a + 1, b[0]);
        wide_raw_row(t + 2, odd, a, b[1]);
#pragma unroll
        for (int i = 2; i < N; i += 2) {
            wide_raw_row(odd + i, t + i, a, b[i]);
            wide_raw_row(t + i + 2, odd + i, a, b[i + 1]);
        }
        // merge: t += odd << 32
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(t[1]) : "r"(odd[0]));
#pragma unroll
        for (int i = 1; i < 2 * N - 2; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(t[i + 1]) : "r"(odd[i]));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[2 * N - 1]));
    }
#endif

#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
    // ONE out-of-line copy of the fully unrolled multiplier per field.  Inlining it at every use makes the hot loops tens
    // of thousands of straight-line instructions long and the kernels instruction-fetch bound (ncu: "no instruction" was
    // the top stall reason of K1 and K8, profiles/r01_*); as a called function the 300 / 600 instructions stay resident
    // in the instruction cache and the operands travel in registers.
    static __device__ __noinline__ mont_t mul_call(mont_t a, mont_t b) {
        mont_t r;
#ifdef ZK_FIELD_MUL_CIOS
        mul_ptx(r.v, a.v, b.v);
#else
        mul_wide(r.v, a.v, b.v);
#endif
        return r;
    }
#endif
    struct pair_t { mont_t a, b; };
#if !defined(ZK_EMU) && !defined(ZK_HOST_ONLY)
    static __device__ __noinline__ pair_t mul2_call(mont_t a0, mont_t b0, mont_t a1, mont_t b1) {
        pair_t r;
        mul_wide2(r.a.v, a0.v, b0.v, r.b.v, a1.v, b1.v);
        return r;
    }
#endif
    // (a0 * b0, a1 * b1)
    static ZK_HD __forceinline__ pair_t mul2(const mont_t &a0, const mont_t &b0, const mont_t &a1, const mont_t &b1) {
#if ZK_FIELD_PTX
        return mul2_call(a0, b0, a1, b1);
#else
        pair_t r;
        r.a = a0 * b0;
        r.b = a1 * b1;
        return r;
#endif
    }
    friend ZK_HD __forceinline__ mont_t operator*(const mont_t &a, const mont_t &b) {
#if ZK_FIELD_PTX
#ifdef ZK_INLINE_FIELD_MUL
        mont_t r;
#ifdef ZK_FIELD_MUL_CIOS
        mul_ptx(r.v, a.v, b.v);
#else
        mul_wide(r.v, a.v, b.v);
#endif
        return r;
#else
        return mul_call(a, b);
#endif
#else
        mont_t r;
        mul_portable(r.v, a.v, b.v);
        return r;
#endif
    }
    ZK_HD __forceinline__ mont_t sqr() const { return *this * *this; }
    ZK_HD __forceinline__ mont_t dbl() const { return *this + *this; }

    // x^e for a canonical little-endian exponent of N limbs (square-and-multiply, MSB first)
    ZK_HD inline mont_t pow_limbs(const uint32_t *e) const {
        mont_t acc = one();
        bool started = false;
        for (int i = N - 1; i >= 0; --i)
            for (int b = 31; b >= 0; --b) {
                if (started) acc = acc.sqr();
                if ((e[i] >> b) & 1) {
                    acc = started ? acc * *this : *this;
                    started = true;
                }
            }
        return acc;
    }
    // multiplicative inverse by Fermat (0 -> 0); kept as the cross-check of inverse()
    ZK_HD inline mont_t inverse_fermat() const { return pow_limbs(C::modm2()); }
    // the same with 4-bit windows: 4 squarings + at most one multiplication per nibble of p - 2 (~32 N squarings + ~8 N multiplications).
    // Tried as the device's inverse() (-DZK_DEVICE_FERMAT_INVERSE) and dropped: normalising one point per thread takes 498 us with it and
    // 521 us with the binary algorithm below, the kernels that invert inside a longer chain got slower (k_g1_to_affine 211 -> 676 us,
    // k_msm_bucket_reduce 515 -> 869 us).  Kept as a third opinion for the self-test.
    ZK_HD inline mont_t inverse_fermat_w4() const {
        mont_t pw[16];
        pw[0] = one();
        pw[1] = *this;
        for (int i = 2; i < 16; ++i) pw[i] = pw[i - 1] * *this;
        const uint32_t *e = C::modm2();
        mont_t acc = one();
        bool started = false;
        for (int i = N - 1; i >= 0; --i)
            for (int b = 28; b >= 0; b -= 4) {
                const uint32_t nib = (e[i] >> b) & 15u;
                if (started) acc = acc.sqr().sqr().sqr().sqr();
                if (nib) {
                    acc = started ? acc * pw[nib] : pw[nib];
                    started = true;
                }
            }
        return acc;
    }

    // ---- raw multi-limb helpers for the binary inversion ---------------------------------------------------------------
    static ZK_HD __forceinline__ bool raw_is_one(const uint32_t *a) {
        uint32_t o = a[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; ++i) o |= a[i];
        return o == 0;
    }
    static ZK_HD __forceinline__ void raw_shr1(uint32_t *a, uint32_t top_in) {   // a = (top_in : a) >> 1
#pragma unroll
        for (int i = 0; i < N - 1; ++i) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
        a[N - 1] = (a[N - 1] >> 1) | (top_in << 31);
    }
    static ZK_HD __forceinline__ uint32_t raw_add(uint32_t *a, const uint32_t *b) {   // a += b, returns the carry out
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            c += (uint64_t) a[i] + b[i];
            a[i] = (uint32_t) c;
            c >>= 32;
        }
        return (uint32_t) c;
    }
    static ZK_HD __forceinline__ void raw_sub(uint32_t *a, const uint32_t *b) {   // a -= b (caller guarantees a >= b or wraps on purpose)
        int64_t bw = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            bw += (int64_t) a[i] - (int64_t) b[i];
            a[i] = (uint32_t) bw;
            bw >>= 32;
        }
    }
    // x = x / 2 mod p
    static ZK_HD __forceinline__ void raw_half_mod(uint32_t *x, const uint32_t *p) {
        uint32_t carry = 0;
        if (x[0] & 1u) carry = raw_add(x, p);
        raw_shr1(x, carry);
    }
    // x = x - y mod p  (x, y < p)
    static ZK_HD __forceinline__ void raw_sub_mod(uint32_t *x, const uint32_t *y, const uint32_t *p) {
        const bool ge = ge_raw(x, y);
        raw_sub(x, y);
        if (!ge) raw_add(x, p);
    }
    // Multiplicative inverse (0 -> 0) by the binary extended Euclidean algorithm on the Montgomery representative
    // (about 2 log2(p) shift/subtract steps instead of the ~1.5 log2(p) field multiplications of Fermat: ~8x fewer
    // instructions for Fp).  Field elements are unique, so the value equals mcl's Fr::inv / Fp::inv.
    ZK_HD inline mont_t inverse() const {
        if (is_zero()) return *this;
#if ZK_ON_DEVICE && defined(ZK_DEVICE_FERMAT_INVERSE)
        return inverse_fermat_w4();   // measured slower on a B200 in every kernel that inverts (one dependent Fp multiplication is ~1 us for a lone warp)
#endif
        uint32_t pm[N], u[N], w[N], x1[N], x2[N];
#pragma unroll
        for (int i = 0; i < N; ++i) { pm[i] = C::mod()[i]; u[i] = v[i]; w[i] = pm[i]; x1[i] = i == 0 ? 1u : 0u; x2[i] = 0u; }
        while (!raw_is_one(u) && !raw_is_one(w)) {
            while (!(u[0] & 1u)) { raw_shr1(u, 0); raw_half_mod(x1, pm); }
            while (!(w[0] & 1u)) { raw_shr1(w, 0); raw_half_mod(x2, pm); }
            if (ge_raw(u, w)) { raw_sub(u, w); raw_sub_mod(x1, x2, pm); }
            else { raw_sub(w, u); raw_sub_mod(x2, x1, pm); }
        }
        // (a R)^-1 = a^-1 R^-1 as a plain integer; one Montgomery multiplication by R^3 brings it to a^-1 R
        mont_t r, r3;
        const bool first = raw_is_one(u);
#pragma unroll
        for (int i = 0; i < N; ++i) { r.v[i] = first ? x1[i] : x2[i]; r3.v[i] = C::r3()[i]; }
        return r * r3;
    }

    // mcl's isNegative(): canonical value >= (p+1)/2  (mcl/include/mcl/fp.hpp:666-671)
    ZK_HD inline bool is_negative() const {
        uint32_t c[N];
        to_canonical(c);
        return ge_raw(c, C::half());
    }
};

typedef mont_t<fr_cfg> fr_t;
typedef mont_t<fp_cfg> fp_t;

// ---- lazy reduction ---------------------------------------------------------------------------------------------------
// A running sum of UNREDUCED products  T = sum_i a_i * b_i  kept as a plain integer of 2N + 1 limbs (one limb of head
// room: 2^32 products).  Only the multiplication half of a Montgomery multiplication is paid per term; the value
// T / R mod p that the sum of the reduced products would have is recovered once, by montgomery_of_wide().  Field
// arithmetic is exact, so the result is the same field element whichever way it is computed.
template <class C> struct lazy_acc_t {
    static constexpr int N = C::N;
    static constexpr int W = 2 * N + 1;
    uint32_t w[W];
    ZK_HD __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < W; ++i) w[i] = 0;
    }
    ZK_HD __forceinline__ void mac(const mont_t<C> &a, const mont_t<C> &b) {
#if ZK_FIELD_PTX
        uint32_t t[2 * N];
        mont_t<C>::mul_raw_wide(t, a.v, b.v);
        asm volatile("add.cc.u32 %0, %0, %1;" : "+r"(w[0]) : "r"(t[0]));
#pragma unroll
        for (int i = 1; i < 2 * N; ++i) asm volatile("addc.cc.u32 %0, %0, %1;" : "+r"(w[i]) : "r"(t[i]));
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(w[2 * N]));
#else
        uint32_t t[2 * N];
#pragma unroll
        for (int i = 0; i < 2 * N; ++i) t[i] = 0;
        for (int i = 0; i < N; ++i) {
            uint64_t c = 0;
            for (int j = 0; j < N; ++j) {
                uint64_t x = (uint64_t) a.v[j] * b.v[i] + t[i + j] + c;
                t[i + j] = (uint32_t) x;
                c = x >> 32;
            }
            t[i + N] = (uint32_t) c;
        }
        uint64_t c = 0;
        for (int i = 0; i < 2 * N; ++i) {
            c += (uint64_t) w[i] + t[i];
            w[i] = (uint32_t) c;
            c >>= 32;
        }
        w[2 * N] += (uint32_t) c;
#endif
    }
};
template <class C> ZK_HD inline mont_t<C> montgomery_of_wide(const uint32_t *c0, const uint32_t *c1, const uint32_t *c2);
// Reduction of a lazy accumulator that holds at most 16 products of REDUCED operands (Fr): T < 16 p^2 < 7.25 p R.
//   U = T >> 256 (9 limbs, < 7.25 p) is brought below p by the ladder 4p, 2p, p;  T' = U' R + (T mod R) < p R, so
//   T' / R = redc(T mod R) + U'  is below 2p: one field addition finishes.  Cost: 64 wide multiply-adds instead of the
//   3 x 128 of montgomery_of_wide; same field element.
ZK_HD __forceinline__ fr_t fr_lazy_reduce_upto16(const lazy_acc_t<fr_cfg> &a) {
#if ZK_FIELD_PTX
    uint32_t u[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) u[i] = a.w[8 + i];
    const uint32_t *mult = ZK_C(fr_MULT9);
#pragma unroll
    for (int k = 2; k >= 0; --k) {
        uint32_t d[9];
        d[0] = ptx::sub_cc(u[0], mult[9 * k]);
#pragma unroll
        for (int i = 1; i < 9; ++i) d[i] = ptx::subc_cc(u[i], mult[9 * k + i]);
        const uint32_t borrow = ptx::subc(0, 0);
#pragma unroll
        for (int i = 0; i < 9; ++i) u[i] = borrow ? u[i] : d[i];
    }
    fr_t lo, hi;
    fr_t::redc_wide(lo.v, a.w);
#pragma unroll
    for (int i = 0; i < 8; ++i) hi.v[i] = u[i];
    return lo + hi;
#else
    uint32_t top[8] = {a.w[16], 0, 0, 0, 0, 0, 0, 0};
    return montgomery_of_wide<fr_cfg>(a.w, a.w + 8, top);
#endif
}

// T / R mod p for a plain integer T = c0 + c1 R + c2 R^2 given as three N-limb chunks:
//   T / R = c0 / R + c1 + c2 R  =  mul(c0, 1) + mul(c1, R mod p) + mul(c2, R^2 mod p)   with mul(x, y) = x y / R.
// The multiplier wants operands below p, so each chunk first loses its multiples of p (R / p < 3 for Fr).
template <class C> ZK_HD inline mont_t<C> montgomery_of_wide(const uint32_t *c0, const uint32_t *c1, const uint32_t *c2) {
    constexpr int N = C::N;
    mont_t<C> x[3], y[3];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        x[0].v[i] = c0[i]; x[1].v[i] = c1[i]; x[2].v[i] = c2[i];
        y[0].v[i] = i == 0 ? 1u : 0u; y[1].v[i] = C::one()[i]; y[2].v[i] = C::r2()[i];
    }
    uint32_t pm[N];
#pragma unroll
    for (int i = 0; i < N; ++i) pm[i] = C::mod()[i];
    for (int k = 0; k < 3; ++k)
        while (mont_t<C>::ge_raw(x[k].v, pm)) mont_t<C>::raw_sub(x[k].v, pm);
    return x[0] * y[0] + x[1] * y[1] + x[2] * y[2];
}
typedef lazy_acc_t<fr_cfg> fr_lazy_t;

static_assert(sizeof(fr_t) == 32, "Fr must match mcl's 32-byte layout");
static_assert(sizeof(fp_t) == 48, "Fp must match mcl's 48-byte layout");

}  // namespace zk
