"""The drop-in boundary: libnrd_b200.so loads on a CPU-only box and exports every symbol include/*.h declares,
with the POD sizes of the reference headers (External/NRD/Include/NRD.h:60-79, NRDDescs.h, NRDSettings.h)."""
import ctypes as C
import os
import re

import pytest

from nrd_sample_b200 import executor, nrd_api as api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, pattern):
    text = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(pattern, text)))


def test_library_exports_every_declared_symbol(product_lib):
    lib = C.CDLL(product_lib)
    nrd_syms = _declared("nrd_b200.h", r"NRD_API [^;(]*?\b(\w+)\(")
    cu_syms = _declared("nrdcu.h", r"NRDCU_API [^;(]*?\b(nrdcu\w+)\(")
    assert set(nrd_syms) == set(api.EXPORTED_SYMBOLS)
    assert set(cu_syms) == set(executor.NRDCU_SYMBOLS)
    for s in nrd_syms + cu_syms:
        assert hasattr(lib, s), f"missing export {s}"


def test_pod_sizes_match_reference_headers():
    # measured by compiling the reference headers with g++ 13.3 (x86-64); static_asserted in include/nrd_b200.h too
    assert C.sizeof(api.CommonSettings) == 432
    assert C.sizeof(api.ReblurSettings) == 120
    assert C.sizeof(api.SigmaSettings) == 20
    assert C.sizeof(api.DispatchDesc) == 56
    assert C.sizeof(api.PipelineDesc) == 320
    assert C.sizeof(api.ResourceDesc) == 12
    assert C.sizeof(api.InstanceDesc) == 112
    assert C.sizeof(api.LibraryDesc) == 40
    assert C.sizeof(api.InstanceCreationDesc) == 48


def test_library_desc_and_strings(host_library):
    d = host_library.library_desc()
    assert (d.versionMajor, d.versionMinor) == (4, 17)
    assert d.normalEncoding == 2 and d.roughnessEncoding == 1  # R10G10B10A2, linear roughness
    sup = host_library.supported_denoisers()
    assert int(api.Denoiser.REBLUR_DIFFUSE_SPECULAR) in sup and int(api.Denoiser.SIGMA_SHADOW) in sup
    assert host_library.lib.GetDenoiserString(int(api.Denoiser.REBLUR_DIFFUSE_SPECULAR)) == b"REBLUR_DIFFUSE_SPECULAR"
    assert host_library.lib.GetResourceTypeString(int(api.ResourceType.IN_VIEWZ)) == b"IN_VIEWZ"


def test_error_behaviour(host_library):
    # a denoiser outside LibraryDesc::supportedDenoisers -> UNSUPPORTED (InstanceImpl.cpp:95-102); all 19 of the reference are supported, so: an id past the enum
    assert sorted(host_library.supported_denoisers()) == list(range(19))
    inst = api.NrdInstance(host_library, [(0, 19)])
    assert inst.result == api.Result.UNSUPPORTED
    # duplicate identifiers -> NON_UNIQUE_IDENTIFIER (:104-108)
    inst = api.NrdInstance(host_library, [(3, api.Denoiser.REBLUR_DIFFUSE_SPECULAR), (3, api.Denoiser.SIGMA_SHADOW)])
    assert inst.result == api.Result.NON_UNIQUE_IDENTIFIER
    inst = api.NrdInstance(host_library, [(1, api.Denoiser.REBLUR_DIFFUSE_SPECULAR)])
    assert inst.result == api.Result.SUCCESS
    # empty identifier list -> SUCCESS with 0 dispatches (:492-497); unknown identifier for settings -> INVALID_ARGUMENT (:484)
    r, d = inst.get_compute_dispatches([])
    assert r == api.Result.SUCCESS and d == []
    assert inst.set_denoiser_settings(99, api.ReblurSettings()) == api.Result.INVALID_ARGUMENT
    # invalid common settings (zero sizes) -> INVALID_ARGUMENT (:279-328, 453)
    assert inst.set_common_settings(api.CommonSettings()) == api.Result.INVALID_ARGUMENT


def test_executor_fails_loudly_without_gpu(product_lib):
    import torch
    if torch.cuda.is_available():
        return
    try:
        executor.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 64, 64)
    except executor.NrdcuError as e:
        assert "no CUDA device" in str(e)
    else:
        raise AssertionError("nrdcuCreate must fail without a CUDA device (no CPU fallback)")


def _build_integration_twin(tmp_path):
    """g++ the NRDSample-shaped call sequence of tests/integration_twin_main.cpp against include/NRDIntegrationCuda.h and the product library."""
    import subprocess
    from nrd_sample_b200 import build
    lib = build.build()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "integration_twin")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
           os.path.join(root, "tests", "integration_twin_main.cpp"), "-o", exe, lib, "-L", os.path.join(cuda, "lib64"), "-lcudart", f"-Wl,-rpath,{os.path.dirname(lib)}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_integration_twin_compiles_and_fails_loudly_without_a_device(tmp_path):
    """include/NRDIntegrationCuda.h mirrors nrd::Integration ( NRDIntegration.h:211-277: Recreate / NewFrame / SetCommonSettings / SetDenoiserSettings /
    Denoise / Destroy + the memory getters ). It must compile warning-free as plain C++17, and with no CUDA device Recreate reports FAILURE — no CPU path."""
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: the gpu-marked twin test covers this box")
    exe = _build_integration_twin(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "no CUDA device" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_integration_twin_denoises_on_the_device(tmp_path):
    import subprocess
    exe = _build_integration_twin(tmp_path)
    r = subprocess.run([exe, "gpu"], capture_output=True, text=True)
    assert r.returncode == 0 and "denoised 2 frames" in r.stdout, r.stdout + r.stderr
    assert "persistent" in r.stdout


def test_every_pipeline_of_every_denoiser_resolves_to_a_kernel(host_library, product_lib):
    """csrc/host/pipeline_key.cpp: each shaderIdentifier an instance can emit ( all 19 denoisers ) maps to a kernel — without a device the call gets as far as checking its
    arguments ( INVALID_ARGUMENT ), an identifier with no kernel stops earlier with UNSUPPORTED. Permutations outside Shaders.cfg are UNSUPPORTED."""
    lib = executor.load()
    seen = set()
    for den in range(19):
        inst = api.NrdInstance(host_library, [(1, den)])
        assert inst.result == api.Result.SUCCESS
        seen.update(inst.shader_identifiers())
    assert len(seen) > 100
    for ident in sorted(seen):
        rc = lib.nrdcuDispatch(ident.encode(), None, 0, None, 0, 0, None)
        assert rc != int(api.Result.UNSUPPORTED) and rc != 0, (ident, rc, lib.nrdcuGetLastError())
    for bogus in ("REBLUR_Blur.cs.hlsl|NRD_SIGNAL=FOO|NRD_MODE=SH", "REBLUR_Blur.cs.hlsl", "RELAX_Atrous.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=OCCLUSION", "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=1",
                  "REBLUR_PrePass.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=SH|MODE_5X5=1", "Unknown.cs.hlsl", "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=2|FIRST_PASS=0"):
        assert lib.nrdcuDispatch(bogus.encode(), None, 0, None, 0, 0, None) == int(api.Result.UNSUPPORTED), bogus


def test_every_chain_kernel_starts_with_the_dependency_wait():
    """launchK ( csrc/kernels/common.cuh ) launches with programmatic stream serialization: a kernel that did not begin with pdlEntry( ) — griddepcontrol.wait —
    could read what its predecessor has not written yet. Every __global__ function of the files that launch through launchK must start with it."""
    import glob
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "nrd_sample_b200", "csrc")
    checked = 0
    for path in sorted(glob.glob(os.path.join(root, "kernels", "*.cu")) + glob.glob(os.path.join(root, "*.cu"))):
        src = open(path).read()
        if "launchK(" not in src:
            continue
        for m in re.finditer(r"__global__", src):
            i, depth = m.end(), 0
            while not (src[i] in "{;" and depth == 0):
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            if src[i] == ";":
                continue
            assert src[i + 1:].lstrip().startswith("pdlEntry();"), f"{os.path.basename(path)}: kernel at offset {m.start()} does not start with pdlEntry()"
            checked += 1
    assert checked >= 30
    for path in glob.glob(os.path.join(root, "kernels", "*.cu")):   # and nothing else goes through launchK
        src = open(path).read()
        if "launchK(" not in src:
            assert "pdlEntry" not in src
