"""Build of the compiled `maxk_kernels` module, the shape of the reference's setup.py:13-40 (one
CUDAExtension named maxk_kernels) -- except that the extension holds no kernels: it is a pybind11 /
ATen front end over the C ABI of ../libmaxk_b200.so, which nvcc builds for compute_100a
(spgemm_gnn_b200/build.py).  In-tree:

    python spgemm_gnn_b200/binding/setup.py build_ext --inplace     # or spgemm_gnn_b200.build.build_binding()
"""
import os

from setuptools import setup
from torch.utils.cpp_extension import BuildExtension, CUDAExtension

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)

setup(
    name="maxk_kernels",
    ext_modules=[
        CUDAExtension(
            name="maxk_kernels_ext",
            sources=[os.path.join(HERE, "maxk_bindings.cpp")],
            include_dirs=[os.path.join(ROOT, "include")],
            library_dirs=[PKG],
            libraries=["maxk_b200"],
            extra_compile_args={"cxx": ["-O2", "-std=c++17"]},
            extra_link_args=["-Wl,-rpath," + PKG],
        )
    ],
    cmdclass={"build_ext": BuildExtension},
)
