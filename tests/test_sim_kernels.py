"""CPU-only: the CUDA kernel sources, compiled with g++ against the SIMT shim (tests/cpu_sim), checked
against the oracle and the reference-generated golden fixtures.  This validates index math, layouts and
the order of floating-point operations; the same assertions run on the real device in test_gpu_parity.py."""
import os

import pytest

import parity_cases as pc
from backends import SimBackend
from oracle import qmf_port as port


@pytest.fixture(scope="module")
def sim():
    return SimBackend()


@pytest.mark.parametrize("shape", [(45, 70), (101, 131), (64, 64), (137, 250)])
def test_frontend_bit_exact(sim, shape):
    pc.check_frontend(sim, port.s_nat(3, *shape))


def test_frontend_other_patches_and_rgb(sim):
    pc.check_frontend(sim, port.s_nat(4, 96, 80), patch=(4, 4))
    pc.check_frontend(sim, port.s_nat(4, 96, 80), patch=(16, 16))
    pc.check_frontend(sim, port.s_nat(4, 50, 70), color_space="RGB")
    pc.check_frontend(sim, port.s_nat(4, 52, 72), color_space="RGB")  # W % 8 == 0: vectorised RGB kernel, rows padded


@pytest.mark.parametrize("name", ["snat7_45x70_q7", "snat8_101x131_q7", "snat9_128x192_q7"])
def test_teacher_forced_factors_bit_exact(sim, manifest, name):
    pc.check_teacher_forced(sim, manifest, name)


# The 256x384 cases take minutes on the shim (768 OS threads per block); they run on the GPU in
# tests/test_gpu_parity.py and here only with LRFB_SIM_FULL=1.
_SIM_BIG = [] if not os.environ.get("LRFB_SIM_FULL") else ["snat1000_256x384_b8", "snat1000_256x384_it2"]


@pytest.mark.parametrize("name", ["snat7_45x70_q7", "snat9_128x192_q7", "snat1000_128x192_rgb"] + _SIM_BIG)
def test_own_svd_init_sign_aligned(sim, manifest, name):
    pc.check_free_running_sign_aligned(sim, manifest, name)


def test_product_path_reproduces_lapack_signs_and_reference_factors(sim, manifest):
    """No hook: the eigen-solver's closed-form sign rule gives LAPACK's signs, so the free-running encode equals the
    reference's factors (128x192: luma M = 384, chroma M = 96 >= 1.5 N)."""
    pc.check_product_signs(sim, manifest, "snat9_128x192_q7")
    pc.check_product_path_identical(sim, manifest, "snat9_128x192_q7")


@pytest.mark.parametrize("kind", ["flat", "black", "half_flat"])
def test_degenerate_images_match_oracle(sim, kind):
    """Rank-one / all-zero planes (s = 0 in U = XV/s, rank-deficient Gram): identical factors."""
    pc.check_degenerate_image(sim, kind)


@pytest.mark.parametrize("name", ["snat7_45x70_q7", "snat8_101x131_q7", "snat1000_256x384_p4",
                                  "snat1000_128x192_rgb"])
def test_decode_and_sse_exact(sim, manifest, name):
    pc.check_decode(sim, manifest, name)
