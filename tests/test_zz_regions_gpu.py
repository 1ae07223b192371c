"""GPU tier of the row regions (tests/test_regions.py holds the CPU tier and the worker): the nvcc-built library with two
processes on ONE device.  The file sorts last on purpose -- it is the only GPU test that spawns processes and talks
over sockets, and `pytest -x` should have run everything else by the time it starts."""
import pytest
from test_regions import _run


@pytest.mark.gpu
def test_gpu_row_regions_two_processes_one_device(ref):
    """the nvcc-built library: two processes share cuda:0 and exchange through the callback transport (gloo, staged
    through host memory) -- the same C++ and kernels as the NCCL path of tools/region_check.py, inside `pytest -m gpu`"""
    _run(2, [0, 1, 2], timeout=240, gpu=True)
