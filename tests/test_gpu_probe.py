"""GPU test: the UMMA descriptor conventions csrc/fa_tc_fwd.cu relies on, checked on one tile."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("D", [64, 128])
def test_umma_descriptor_conventions(D, dtype):
    import probe_umma
    box = 64 * D * 2
    assert probe_umma.run(0, D, dtype, box, 1024, 2048, 0) < 1e-5     # S = Q K^T, MN-major SW128
    assert probe_umma.run(1, D, dtype, 16, 1024, 32, 4) < 1e-5        # O = P V,  P in TMEM, V K-major SW128
