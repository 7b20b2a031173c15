"""Test-only helper: compile a CUDA snippet with the engine's NVRTC binding and launch it through
cuda-python's driver API (so tests can exercise device helpers of csrc/inflx_device.cuh alone)."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def device_header() -> str:
    with open(os.path.join(ROOT, "inflatox_b200", "csrc", "inflx_device.cuh")) as fh:
        return fh.read()


class Module:
    def __init__(self, source: str, fmad: bool = False):
        from cuda.bindings import driver as cu

        from inflatox_b200.compiler import nvrtc_compile

        self.cu = cu
        opts = ["--gpu-architecture=sm_100a", "--std=c++17", f"--fmad={'true' if fmad else 'false'}",
                "--prec-div=true", "--prec-sqrt=true"]
        cubin = nvrtc_compile(device_header() + source, "test_kernels.cu", opts)
        (err,) = cu.cuInit(0)
        assert err == cu.CUresult.CUDA_SUCCESS, err
        err, dev = cu.cuDeviceGet(0)
        err, self.ctx = cu.cuDevicePrimaryCtxRetain(dev)
        (err,) = cu.cuCtxSetCurrent(self.ctx)
        err, self.mod = cu.cuModuleLoadData(cubin)
        assert err == cu.CUresult.CUDA_SUCCESS, err

    def launch(self, name: str, n: int, arrays: list[np.ndarray], outputs: list[np.ndarray]):
        """kernel(in..., out..., int n): 1-D launch over n elements."""
        cu = self.cu
        err, fn = cu.cuModuleGetFunction(self.mod, name.encode())
        assert err == cu.CUresult.CUDA_SUCCESS, (name, err)
        ptrs = []
        for a in arrays + outputs:
            err, d = cu.cuMemAlloc(a.nbytes)
            assert err == cu.CUresult.CUDA_SUCCESS, err
            ptrs.append(d)
        for a, d in zip(arrays, ptrs):
            (err,) = cu.cuMemcpyHtoD(d, a.ctypes.data, a.nbytes)
        args = [np.array([int(d)], dtype=np.uint64) for d in ptrs] + [np.array([n], dtype=np.int32)]
        argp = np.array([a.ctypes.data for a in args], dtype=np.uint64)
        (err,) = cu.cuLaunchKernel(fn, (n + 255) // 256, 1, 1, 256, 1, 1, 0, 0, argp.ctypes.data, 0)
        assert err == cu.CUresult.CUDA_SUCCESS, err
        (err,) = cu.cuCtxSynchronize()
        assert err == cu.CUresult.CUDA_SUCCESS, err
        for a, d in zip(outputs, ptrs[len(arrays):]):
            (err,) = cu.cuMemcpyDtoH(a.ctypes.data, d, a.nbytes)
        for d in ptrs:
            cu.cuMemFree(d)
