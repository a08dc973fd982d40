"""TEST INFRASTRUCTURE -- builds the reference's OWN CUDA op into oracle/_ref/ (git-ignored, travels to the GPU box).

    python oracle/build_ref_cu.py          (needs /root/reference; no GPU: nvcc cross-compiles sm_100a)

The reference's only native code is splat/c/render.cu, JIT-built by its scene through
`load_cuda(cuda_src, cpp_src, ["render_image"])` (splat/utils.py:426-434, splat/gaussian_scene.py:240-261).  This
script does exactly that build -- same torch.utils.cpp_extension.load_inline call, same declaration string, same -O1
-- from the source where it lies under /root/reference, and keeps only the resulting extension module
(oracle/_ref/gsb_ref_render_cu.so).  No reference source is copied into the repository.  tests/test_gpu_ref_cu.py
loads the module on the GPU box and pins `semantics = GSB_SEM_REF_CU` to what the reference kernel itself computes.
"""
import glob
import os
import shutil
import sys
import tempfile

REF = os.environ.get("GSB_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "gsb_ref_render_cu"

# the declaration compile_cuda_ext passes as cpp_sources (splat/gaussian_scene.py:244-257): the op's signature
CPP_DECL = """
torch::Tensor render_image(int image_height, int image_width, int tile_size, torch::Tensor point_means,
    torch::Tensor point_colors, torch::Tensor inverse_covariance_2d, torch::Tensor min_x, torch::Tensor max_x,
    torch::Tensor min_y, torch::Tensor max_y, torch::Tensor opacity);
"""


def build() -> str:
    src = os.path.join(REF, "splat", "c", "render.cu")
    if not os.path.exists(src):
        raise RuntimeError(f"{src} not found: the reference checkout is needed to build its op")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load_inline

    tmp = tempfile.mkdtemp(prefix="gsb_ref_cu_")
    load_inline(name=NAME, cpp_sources=[CPP_DECL], cuda_sources=[open(src).read()], functions=["render_image"],
                extra_cuda_cflags=["-O1"], build_directory=tmp, verbose=False, is_python_module=False)
    so = glob.glob(os.path.join(tmp, NAME + "*.so"))
    if not so:
        raise RuntimeError("the extension did not build")
    os.makedirs(OUT_DIR, exist_ok=True)
    dst = os.path.join(OUT_DIR, NAME + ".so")
    shutil.copyfile(so[0], dst)
    shutil.rmtree(tmp, ignore_errors=True)  # generated copies of the source stay out of the tree
    return dst


if __name__ == "__main__":
    print(build())
    sys.exit(0)
