"""GPU parity on mixed batches whose query and target lengths sit on the 32/64-bit word boundaries (cases.boundary_mix_cases:
alphabets of 1..256 symbols, every mode, task and bound, shared and per-query targets), through the product's C ABI against
the reference build.  The same generator runs on the CPU emulation in test_engine_emul.py and in scripts/stress.py."""
import pytest

import cases
import parity
from helpers import product

pytestmark = pytest.mark.gpu


def test_lengths_on_word_boundaries_in_mixed_batches():
    lib = product()
    assert lib.lib.edlibB200Available() == 1, "CUDA path unavailable: the product has no CPU fallback"
    assert parity.run_batches(lib, 7, 60, gen=cases.boundary_mix_cases) > 3000
