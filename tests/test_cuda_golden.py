"""GPU: the CUDA path (through the C ABI) must reproduce the reference's golden vectors bit-exactly."""
import pytest

import golden_util as gu

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("logic_variant")]


@pytest.mark.parametrize("name", gu.fixture_names())
def test_cuda_matches_reference_golden(name):
    fx = gu.load(name)
    gu.replay(gu.CudaBackend(fx), fx, check_features=True)


@pytest.mark.parametrize("name", ["cleanup_n2", "cleanup_n5_short_horizon", "harvest_n4"])
def test_cuda_matches_golden_padded_obs(name):
    fx = gu.load(name)
    gu.replay(gu.CudaBackend(fx, num_envs=4, index=3, padded_obs=True), fx, check_features=False)
