"""ctypes binding of the two C interfaces: include/zkcnn_b200.h (the drop-in boundary, CUDA) and include/zkcnn_host.h
(stand-alone host side).  Plain pointers and sizes only; field elements travel as numpy uint64 arrays of shape [n, 4]
(mcl's in-memory Montgomery form), points as [n, 18] (Jacobian x, y, z)."""
import ctypes as C
import os

import numpy as np

R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
R_MONT = (1 << 256) % R_MOD      # Montgomery radix of Fr
P_MONT = (1 << 384) % P_MOD      # Montgomery radix of Fp
R_MONT_INV = pow(R_MONT, -1, R_MOD)
P_MONT_INV = pow(P_MONT, -1, P_MOD)

_u64p = C.POINTER(C.c_uint64)


def _ptr(a):
    return a.ctypes.data_as(_u64p) if a is not None else None


# ---- conversions between Python integers and the in-memory forms --------------------------------------------------------
def fr_to_words(values):
    """canonical integers (any sign) -> [n, 4] uint64 Montgomery limbs"""
    out = np.empty((len(values), 4), dtype=np.uint64)
    for i, v in enumerate(values):
        m = (int(v) % R_MOD) * R_MONT % R_MOD
        for k in range(4):
            out[i, k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def fr_from_words(words):
    w = np.asarray(words, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(w[i, k]) << (64 * k) for k in range(4)) * R_MONT_INV % R_MOD for i in range(len(w))]


def fp_int_from_words(w6):
    return sum(int(w6[k]) << (64 * k) for k in range(6)) * P_MONT_INV % P_MOD


def g1_to_words(points):
    """affine points [(x, y) or None] -> [n, 18] uint64 Jacobian Montgomery (z = 1; None -> all zero)"""
    out = np.zeros((len(points), 18), dtype=np.uint64)
    for i, pt in enumerate(points):
        if pt is None:
            continue
        for c, v in enumerate((pt[0], pt[1], 1)):
            m = v % P_MOD * P_MONT % P_MOD
            for k in range(6):
                out[i, 6 * c + k] = (m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def g1_from_words(words):
    """[n, 18] -> affine (x, y) integers or None for infinity (any z)"""
    w = np.asarray(words, dtype=np.uint64).reshape(-1, 18)
    out = []
    for i in range(len(w)):
        x, y, z = (fp_int_from_words(w[i, 6 * c:6 * c + 6]) for c in range(3))
        if z == 0:
            out.append(None)
            continue
        zi = pow(z, -1, P_MOD)
        out.append((x * zi * zi % P_MOD, y * zi * zi * zi % P_MOD))
    return out


class ZkError(RuntimeError):
    pass


class Lib:
    """libzkcnn_b200.so (or, in the CPU-only test suite, the emulator build of the same sources)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise ZkError(f"{path} is missing: build it with `make lib` (zkcnn_b200 has no CPU fallback)")
        self.path = path
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        d = self.dll
        d.zk_last_error.restype = C.c_char_p
        d.zk_version.restype = C.c_char_p
        d.zk_ctx_create.restype = C.c_void_p
        d.zk_ctx_create.argtypes = [C.c_int]
        d.zk_ctx_destroy.argtypes = [C.c_void_p]
        d.zk_ctx_launch_count.restype = C.c_uint64
        d.zk_ctx_launch_count.argtypes = [C.c_void_p]
        d.zk_profile_enable.argtypes = [C.c_void_p, C.c_int]
        d.zk_set_tunable.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        d.zk_profile_msm_ops.argtypes = [C.c_void_p, _u64p]
        d.zk_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), _u64p, _u64p]
        d.zk_fr_vec_op.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, C.c_uint64]
        d.zk_beta_table.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p, _u64p]
        d.zk_phi_table.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint32, C.c_int, _u64p]
        d.zk_fold_rounds.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint32, C.c_uint64, _u64p, C.c_uint32, _u64p]
        d.zk_fold_rounds2.argtypes = [C.c_void_p, _u64p, _u64p, C.c_int32, C.c_uint64, _u64p, _u64p, C.c_int32, C.c_uint64, _u64p, C.c_uint32, _u64p]
        d.zk_mle_eval.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p, C.c_uint32, _u64p]
        d.zk_cubic_rounds.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p, C.c_uint64, _u64p, C.c_uint64, C.c_uint32, _u64p, C.c_uint32, _u64p]
        d.zk_msm.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64, C.c_uint32, _u64p]
        d.zk_g1_vec_op.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, C.c_uint64]
        d.zk_g1_fixed_base_mul.argtypes = [C.c_void_p, _u64p, _u64p, C.c_uint64, _u64p]
        d.zk_selftest.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
        d.zk_bench_fold.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_float)]
        d.zk_bench_cubic.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        d.zk_bench_field_mul.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
        d.zk_bench_msm.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.POINTER(C.c_float)]
        d.zk_poly_create.argtypes = [C.c_void_p, _u64p, C.c_uint64, _u64p, C.c_uint32]
        d.zk_poly_commit.argtypes = [C.c_void_p, _u64p, C.c_uint32]
        d.zk_poly_evaluate.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p]
        d.zk_poly_init_bullet_prove.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p, C.c_uint32]
        d.zk_poly_bullet_prove.argtypes = [C.c_void_p, _u64p, _u64p, _u64p, _u64p]
        d.zk_poly_bullet_update.argtypes = [C.c_void_p, _u64p]
        d.zk_poly_bullet_prove_all.argtypes = [C.c_void_p, _u64p, C.c_uint32, _u64p, _u64p, _u64p, _u64p]
        d.zk_poly_bullet_open.argtypes = [C.c_void_p, _u64p]

    def version(self):
        return self.dll.zk_version().decode()

    def device_count(self):
        return self.dll.zk_device_count()

    def last_error(self):
        return self.dll.zk_last_error().decode()


PROF_CLASSES = ("fold", "gates", "msm", "tables", "dense", "other", "fold_small")


class Context:
    """One zk_ctx: the stateless primitives and micro-benchmarks of include/zkcnn_b200.h."""

    def __init__(self, lib, device=0):
        self.lib = lib
        self.h = lib.dll.zk_ctx_create(device)
        if not self.h:
            raise ZkError("zk_ctx_create: " + lib.last_error())

    def close(self):
        if self.h:
            self.lib.dll.zk_ctx_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise ZkError(f"{what}: {self.lib.last_error()}")

    def set_tunable(self, name, value):
        self._check(self.lib.dll.zk_set_tunable(self.h, name.encode(), int(value)), "zk_set_tunable")

    def launches(self):
        return int(self.lib.dll.zk_ctx_launch_count(self.h))

    def selftest(self, seed=1, n=1 << 16):
        self._check(self.lib.dll.zk_selftest(self.h, seed, n), "zk_selftest")

    def fr_vec_op(self, op, a, b):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty_like(a)
        self._check(self.lib.dll.zk_fr_vec_op(self.h, op, _ptr(a), _ptr(b), _ptr(out), len(a)), "zk_fr_vec_op")
        return out

    def beta_table(self, r, init):
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        init = np.ascontiguousarray(init, dtype=np.uint64).reshape(4)
        out = np.empty((1 << len(r), 4), dtype=np.uint64)
        self._check(self.lib.dll.zk_beta_table(self.h, _ptr(r) if len(r) else None, len(r), _ptr(init), _ptr(out)), "zk_beta_table")
        return out

    def phi_table(self, rx, scale, n, is_ifft):
        rx = np.ascontiguousarray(rx, dtype=np.uint64).reshape(-1, 4)
        scale = np.ascontiguousarray(scale, dtype=np.uint64).reshape(4)
        out = np.empty((1 << n if is_ifft else 1 << (n - 1), 4), dtype=np.uint64)
        self._check(self.lib.dll.zk_phi_table(self.h, _ptr(rx), _ptr(scale), n, int(is_ifft), _ptr(out)), "zk_phi_table")
        return out

    def fold_rounds(self, V, M, bits, r, n_rounds):
        V = np.ascontiguousarray(V, dtype=np.uint64).reshape(-1, 4)
        M = np.ascontiguousarray(M, dtype=np.uint64).reshape(-1, 4)
        assert len(V) == len(M)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        out = np.empty((n_rounds * 3, 4), dtype=np.uint64)
        self._check(self.lib.dll.zk_fold_rounds(self.h, _ptr(V), _ptr(M), bits, len(V), _ptr(r) if len(r) else None, n_rounds, _ptr(out)),
                    "zk_fold_rounds")
        return out.reshape(n_rounds, 3, 4)

    def fold_rounds2(self, V0, M0, bits0, V1, M1, bits1, r, n_rounds):
        """K1 on two table pairs in lock step (prover::sumcheckUpdate); bits0 < 0: pair 0 absent"""
        arrs = [None if a is None else np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4) for a in (V0, M0, V1, M1)]
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        out = np.empty((n_rounds * 3, 4), dtype=np.uint64)
        n0 = 0 if arrs[0] is None else len(arrs[0])
        n1 = 0 if arrs[2] is None else len(arrs[2])
        self._check(self.lib.dll.zk_fold_rounds2(self.h, _ptr(arrs[0]), _ptr(arrs[1]), bits0, n0, _ptr(arrs[2]), _ptr(arrs[3]), bits1, n1,
                                                 _ptr(r) if len(r) else None, n_rounds, _ptr(out)), "zk_fold_rounds2")
        return out.reshape(n_rounds, 3, 4)

    def mle_eval(self, values, r):
        values = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        out = np.empty(4, dtype=np.uint64)
        self._check(self.lib.dll.zk_mle_eval(self.h, _ptr(values), len(values), _ptr(r) if len(r) else None, len(r), _ptr(out)), "zk_mle_eval")
        return out

    def cubic_rounds(self, mult, V0, V1, bits, r, n_rounds):
        """K2: n_rounds of sumcheckDotProdUpdate1 on stand-alone tables (len(mult) a power of two; len(V0) <= len(V1) <= 2^bits)"""
        mult = np.ascontiguousarray(mult, dtype=np.uint64).reshape(-1, 4)
        V0 = np.ascontiguousarray(V0, dtype=np.uint64).reshape(-1, 4)
        V1 = np.ascontiguousarray(V1, dtype=np.uint64).reshape(-1, 4)
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(-1, 4)
        m_bits = len(mult).bit_length() - 1
        assert len(mult) == 1 << m_bits
        out = np.empty((n_rounds * 4, 4), dtype=np.uint64)
        self._check(self.lib.dll.zk_cubic_rounds(self.h, _ptr(mult), m_bits, _ptr(V0), len(V0), _ptr(V1), len(V1), bits, _ptr(r) if len(r) else None, n_rounds,
                                                 _ptr(out)), "zk_cubic_rounds")
        return out.reshape(n_rounds, 4, 4)

    def msm(self, bases, scalars, n_rows=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 18)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        n = len(bases)
        assert len(scalars) == n * n_rows
        out = np.empty((n_rows, 18), dtype=np.uint64)
        self._check(self.lib.dll.zk_msm(self.h, _ptr(bases), _ptr(scalars), n, n_rows, _ptr(out)), "zk_msm")
        return out

    def g1_vec_op(self, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 18)
        if b is not None:
            b = np.ascontiguousarray(b, dtype=np.uint64)
        out = np.empty_like(a)
        self._check(self.lib.dll.zk_g1_vec_op(self.h, op, _ptr(a), _ptr(b), _ptr(out), len(a)), "zk_g1_vec_op")
        return out

    def g1_fixed_base_mul(self, base, scalars):
        base = np.ascontiguousarray(base, dtype=np.uint64).reshape(18)
        scalars = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.empty((len(scalars), 18), dtype=np.uint64)
        self._check(self.lib.dll.zk_g1_fixed_base_mul(self.h, _ptr(base), _ptr(scalars), len(scalars), _ptr(out)), "zk_g1_fixed_base_mul")
        return out

    def bench_fold(self, bits, iters=20, fold=True):
        ms = C.c_float(0)
        self._check(self.lib.dll.zk_bench_fold(self.h, bits, iters, int(fold), C.byref(ms)), "zk_bench_fold")
        return ms.value

    def bench_cubic(self, bits, m_bits=7, iters=10, live0_bits=None):
        ms = C.c_float(0)
        self._check(self.lib.dll.zk_bench_cubic(self.h, bits, bits if live0_bits is None else live0_bits, m_bits, iters, C.byref(ms)), "zk_bench_cubic")
        return ms.value

    def bench_fp_mul(self, fp=True):
        g = C.c_float(0)
        self._check(self.lib.dll.zk_bench_field_mul(self.h, int(fp), C.byref(g)), "zk_bench_field_mul")
        return g.value

    def bench_msm(self, log_rows, log_cols, scalar_mix=2, iters=3):
        ms = C.c_float(0)
        self._check(self.lib.dll.zk_bench_msm(self.h, log_rows, log_cols, scalar_mix, iters, C.byref(ms)), "zk_bench_msm")
        return ms.value

    # ---- stand-alone Hyrax polynomial (class polyProver) ----
    def poly_create(self, Z, gens):
        Z = np.ascontiguousarray(Z, dtype=np.uint64).reshape(-1, 4)
        gens = np.ascontiguousarray(gens, dtype=np.uint64).reshape(-1, 18)
        self._check(self.lib.dll.zk_poly_create(self.h, _ptr(Z), len(Z), _ptr(gens), len(gens)), "zk_poly_create")

    def poly_commit(self, n_out):
        out = np.empty((n_out, 18), dtype=np.uint64)
        self._check(self.lib.dll.zk_poly_commit(self.h, _ptr(out), n_out), "zk_poly_commit")
        return out

    def poly_evaluate(self, x):
        x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
        out = np.empty(4, dtype=np.uint64)
        self._check(self.lib.dll.zk_poly_evaluate(self.h, _ptr(x), len(x), _ptr(out)), "zk_poly_evaluate")
        return out

    def poly_init_bullet_prove(self, lx, rx):
        lx = np.ascontiguousarray(lx, dtype=np.uint64).reshape(-1, 4)
        rx = np.ascontiguousarray(rx, dtype=np.uint64).reshape(-1, 4)
        self._check(self.lib.dll.zk_poly_init_bullet_prove(self.h, _ptr(lx) if len(lx) else None, len(lx), _ptr(rx) if len(rx) else None, len(rx)),
                    "zk_poly_init_bullet_prove")

    def poly_bullet_prove(self):
        lc, rc = np.empty(18, dtype=np.uint64), np.empty(18, dtype=np.uint64)
        ly, ry = np.empty(4, dtype=np.uint64), np.empty(4, dtype=np.uint64)
        self._check(self.lib.dll.zk_poly_bullet_prove(self.h, _ptr(lc), _ptr(rc), _ptr(ly), _ptr(ry)), "zk_poly_bullet_prove")
        return lc, rc, ly, ry

    def poly_bullet_update(self, r):
        r = np.ascontiguousarray(r, dtype=np.uint64).reshape(4)
        self._check(self.lib.dll.zk_poly_bullet_update(self.h, _ptr(r)), "zk_poly_bullet_update")

    def poly_bullet_prove_all(self, rands):
        """every remaining round at once; returns (lcomm[n][18], rcomm[n][18], ly[n][4], ry[n][4])"""
        rands = np.ascontiguousarray(rands, dtype=np.uint64).reshape(-1, 4)
        n = len(rands)
        lc, rc = np.empty((n, 18), dtype=np.uint64), np.empty((n, 18), dtype=np.uint64)
        ly, ry = np.empty((n, 4), dtype=np.uint64), np.empty((n, 4), dtype=np.uint64)
        self._check(self.lib.dll.zk_poly_bullet_prove_all(self.h, _ptr(rands), n, _ptr(lc), _ptr(rc), _ptr(ly), _ptr(ry)), "zk_poly_bullet_prove_all")
        return lc, rc, ly, ry

    def poly_bullet_open(self):
        out = np.empty(4, dtype=np.uint64)
        self._check(self.lib.dll.zk_poly_bullet_open(self.h, _ptr(out)), "zk_poly_bullet_open")
        return out


class Stats(C.Structure):
    _fields_ = [("ok", C.c_int32), ("n_layers", C.c_uint32), ("input_size", C.c_uint64), ("n_fr", C.c_uint64), ("n_g1", C.c_uint64),
                ("proof_bytes", C.c_uint64), ("fnv1a", C.c_uint64), ("challenges", C.c_uint64), ("gpu_launches", C.c_uint64),
                ("prove_s", C.c_double), ("poly_s", C.c_double), ("upload_s", C.c_double), ("wall_s", C.c_double),
                ("verifier_s", C.c_double), ("gkr_kb", C.c_double), ("poly_kb", C.c_double), ("h2d_bytes", C.c_uint64), ("checks", C.c_uint32),
                ("witness_path", C.c_uint32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


REAL_GENERATORS, CHECK_PREDICATES, WITNESS_RESIDENT, FIXED_GENERATORS, ROUND_BY_ROUND, PREFETCH_NEXT, NO_HASH = 1, 2, 4, 8, 16, 32, 64
PROVER_ONLY, CSPRNG_CHALLENGES, FIAT_SHAMIR, HOST_PREDICATES = 128, 256, 512, 1024
CHECKED_ROUND_SUMS, CHECKED_PREDICATES, CHECKED_INPUT_GR, CHECKED_G1 = 1, 2, 4, 8
CHECKED_ALL = 15


class HostLib:
    """libzkcnn_host.so: model zoo, circuit compiler + witness generator, protocol driver (include/zkcnn_host.h)."""

    def __init__(self, path):
        if not os.path.exists(path):
            raise ZkError(f"{path} is missing: build it with `make host`")
        self.dll = C.CDLL(path, mode=C.RTLD_LOCAL)
        d = self.dll
        d.zkh_last_error.restype = C.c_char_p
        d.zkh_create.restype = C.c_void_p
        d.zkh_create.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int]
        d.zkh_destroy.argtypes = [C.c_void_p]
        d.zkh_input_count.restype = C.c_int64
        d.zkh_input_count.argtypes = [C.c_void_p]
        d.zkh_input_file.argtypes = [C.c_void_p, C.c_char_p]
        d.zkh_parse_numbers.restype = C.c_int64
        d.zkh_parse_numbers.argtypes = [C.c_char_p, C.POINTER(C.c_double), C.c_uint64]
        d.zkh_input_values.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64]
        d.zkh_build.argtypes = [C.c_void_p]
        d.zkh_prefetch_witness.argtypes = [C.c_void_p]
        d.zkh_prove.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(Stats)]
        d.zkh_prove_image.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(Stats)]
        d.zkh_set_image.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint64]
        d.zkh_proof.restype = C.POINTER(C.c_uint8)
        d.zkh_proof.argtypes = [C.c_void_p, _u64p]
        d.zkh_inferred_class.argtypes = [C.c_void_p, C.c_int]
        d.zkh_circuit_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        d.zkh_table_dump.argtypes = [C.c_void_p, C.c_char_p]
        d.zkh_context.restype = C.c_void_p
        d.zkh_context.argtypes = [C.c_void_p]

    def last_error(self):
        return self.dll.zkh_last_error().decode()

    def parse_numbers(self, path):
        """the reference's text input format -> float64 array (block-wise strtod; the reference re-parses with `ifstream >> double` in every run)"""
        n = self.dll.zkh_parse_numbers(str(path).encode(), None, 0)
        if n < 0:
            raise ZkError("zkh_parse_numbers: " + self.last_error())
        out = np.empty(n, dtype=np.float64)
        if self.dll.zkh_parse_numbers(str(path).encode(), out.ctypes.data_as(C.POINTER(C.c_double)), n) != n:
            raise ZkError("zkh_parse_numbers: file changed while reading")
        return out


class Session:
    """One model on one GPU: build the circuit and the witness once, then prove."""

    def __init__(self, hostlib, model, network="", pic_cnt=1, device=0):
        self.lib = hostlib
        self.h = hostlib.dll.zkh_create(model.encode(), network.encode(), pic_cnt, device)
        if not self.h:
            raise ZkError("zkh_create: " + hostlib.last_error())

    def close(self):
        if self.h:
            self.lib.dll.zkh_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise ZkError(f"{what}: {self.lib.last_error()}")

    def input_count(self):
        return int(self.lib.dll.zkh_input_count(self.h))

    def input_file(self, path):
        self._check(self.lib.dll.zkh_input_file(self.h, str(path).encode()), "zkh_input_file")

    def input_values(self, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        self._check(self.lib.dll.zkh_input_values(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), len(v)), "zkh_input_values")

    def build(self):
        self._check(self.lib.dll.zkh_build(self.h), "zkh_build")

    def prefetch_witness(self):
        self._check(self.lib.dll.zkh_prefetch_witness(self.h), "zkh_prefetch_witness")

    def prove(self, seed, flags=0):
        st = Stats()
        self._check(self.lib.dll.zkh_prove(self.h, seed, flags, C.byref(st)), "zkh_prove")
        return st.as_dict()

    def prove_image(self, pixels, seed, flags=0):
        """a new picture for the built model: witness regenerated on the device (stats["witness_path"] == 1) or, if the picture does not
        fit the circuit's quantisation decisions, circuit + witness rebuilt on the host (== 2)"""
        v = np.ascontiguousarray(pixels, dtype=np.float64)
        st = Stats()
        self._check(self.lib.dll.zkh_prove_image(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), len(v), seed, flags, C.byref(st)), "zkh_prove_image")
        return st.as_dict()

    def set_image(self, pixels):
        """the witness part of prove_image; returns 1 (regenerated on the device) or 2 (circuit + witness rebuilt on the host)"""
        v = np.ascontiguousarray(pixels, dtype=np.float64)
        rc = self.lib.dll.zkh_set_image(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), len(v))
        if rc < 0:
            raise ZkError("zkh_set_image: " + self.lib.last_error())
        return rc

    def proof(self):
        n = C.c_uint64(0)
        p = self.lib.dll.zkh_proof(self.h, C.byref(n))
        return bytes(C.cast(p, C.POINTER(C.c_uint8 * n.value)).contents) if n.value else b""

    def inferred_class(self, picture=0):
        return self.lib.dll.zkh_inferred_class(self.h, picture)

    def circuit_dump(self, path, with_hashes=True):
        self._check(self.lib.dll.zkh_circuit_dump(self.h, str(path).encode(), int(with_hashes)), "zkh_circuit_dump")

    def table_dump(self, path):
        """test hook: hashes of the bookkeeping tables after every Init* call of the following proofs -> path (None: off)"""
        self._check(self.lib.dll.zkh_table_dump(self.h, str(path).encode() if path is not None else None), "zkh_table_dump")

    def context_handle(self):
        """the zk_ctx of this session's prover (for zk_profile_*); None before the first proof"""
        return self.lib.dll.zkh_context(self.h)
