"""CPU suite: the CUDA kernel sources compiled for the test-only emulator, driven through the same C ABI."""
import _cases as cases


def test_fr_vec_ops(emu_lib):
    cases.case_fr_vec_ops(emu_lib)


def test_selftest(emu_lib):
    """the device self-test's identities on the host build of the same arithmetic (multiplier variants, lazy sums, the three inversions)"""
    from zkcnn_b200._binding import Context
    with Context(emu_lib) as ctx:
        ctx.selftest(seed=7, n=1024)


def test_fr_kat(emu_lib, kat):
    cases.case_fr_kat(emu_lib, kat)


def test_beta_tables(emu_lib, kat):
    cases.case_beta_tables(emu_lib, kat)


def test_phi_tables(emu_lib, kat):
    cases.case_phi_tables(emu_lib, kat)


def test_fold_rounds(emu_lib):
    cases.case_fold_rounds(emu_lib)


def test_cubic_rounds(emu_lib):
    small = ((1, 0, 2, 2), (3, 1, 5, 8), (5, 2, 12, 29), (8, 3, 100, 250), (10, 4, 301, 1024))
    cases.case_cubic_rounds(emu_lib, shapes=small)                                            # default selection
    cases.case_cubic_rounds(emu_lib, shapes=small, tunables={"cubic_factored_min_iters": 1})         # factored form everywhere
    cases.case_cubic_rounds(emu_lib, shapes=small, tunables={"cubic_factored_min_iters": 1 << 30})   # direct form everywhere
    # several iterations per thread with a multiplier period of more than one CTA: the grid must stay a multiple of the period
    multi = ((11, 10, 1500, 2048), (12, 9, 2500, 4000))
    cases.case_cubic_rounds(emu_lib, shapes=multi, tunables={"cubic_max_grid": 5})
    cases.case_cubic_rounds(emu_lib, shapes=multi, tunables={"cubic_max_grid": 3, "cubic_factored_min_iters": 1 << 30})


def test_fold_rounds_two_pairs_and_fused_tail(emu_lib, kat):
    small = ((3, 5, 2, 4), (6, 40, 4, 9), (9, 400, 11, 2048), (11, 1500, 5, 32))
    cases.case_fold_rounds_two_pairs(emu_lib, shapes=small)
    cases.case_fold_rounds_two_pairs(emu_lib, shapes=small, tunables={"unit_batch": 1})                         # fused tail from 1024 entries down
    cases.case_fold_rounds_two_pairs(emu_lib, shapes=small, tunables={"unit_batch": 1, "tail_max_entries": 64})
    cases.case_fold_rounds_two_pairs(emu_lib, shapes=small, tunables={"unit_batch": 1, "tail": 0})
    cases.case_round_kats(emu_lib, kat, tunables={"unit_batch": 1})


def test_round_kats_from_the_reference(emu_lib, kat):
    cases.case_round_kats(emu_lib, kat)
    cases.case_round_kats(emu_lib, kat, tunables={"thin_max_pairs": 0, "cubic_factored_min_iters": 1})


def test_g1_ops(emu_lib, kat):
    cases.case_g1_ops(emu_lib, kat)


def test_msm(emu_lib, kat):
    cases.case_msm(emu_lib, kat)


def test_hyrax_kat(emu_lib, kat):
    cases.case_hyrax_kat(emu_lib, kat)


def test_hyrax_vs_port(emu_lib):
    cases.case_hyrax_vs_port(emu_lib, bl=5)


def test_msm_many_rows(emu_lib):
    cases.case_msm_many_rows(emu_lib, n=40, rows=17)


def test_msm_bucket_shapes(emu_lib):
    cases.case_msm_bucket_shapes(emu_lib)


def test_fixed_base_mul(emu_lib, kat):
    cases.case_fixed_base_mul(emu_lib, kat)


def test_hyrax_vs_port_small_path(emu_lib):
    cases.case_hyrax_vs_port(emu_lib, bl=8, seed=1001)     # 16 commitment rows: the small-multiples path
