"""The C-ABI library loads without a GPU and exports every entry point include/dicow_b200.h declares; the ctypes struct
mirrors agree with the header field by field (names, order); no compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dicow_b200.h")


@pytest.fixture(scope="module")
def lib():
    from ts_asr_whisper_b200 import build, lib as _lib
    build.build()
    return _lib


def _header_text():
    with open(HEADER) as f:
        return re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)


def test_exports_match_header(lib):
    declared = re.findall(r"DICOW_API\s+[\w\s\*]+?\b(dicow_\w+)\s*\(", _header_text())
    assert declared, "no declarations parsed"
    assert sorted(declared) == sorted(lib.EXPORTED_SYMBOLS)
    so = lib.load_library()
    for name in declared:
        assert hasattr(so, name), f"{name} not exported by libdicow_b200.so"
    from ts_asr_whisper_b200 import lib as _l
    assert so.dicow_abi_version() == _l.ABI_VERSION


@pytest.mark.parametrize("cname,pyname", [("dicow_gemm_args_t", "GemmArgs"), ("dicow_fddt_ln_args_t", "FddtLnArgs"),
                                          ("dicow_attention_args_t", "AttentionArgs"), ("dicow_logmel_args_t", "LogmelArgs"),
                                          ("dicow_attention_bwd_args_t", "AttentionBwdArgs"), ("dicow_ln_bwd_args_t", "LnBwdArgs"),
                                          ("dicow_ctc_bwd_args_t", "CtcBwdArgs"),
                                          ("dicow_softlabel_ce_bwd_args_t", "SoftlabelCeBwdArgs"),
                                          ("dicow_gemm_skinny_args_t", "GemmSkinnyArgs"),
                                          ("dicow_decode_attention_args_t", "DecodeAttentionArgs"),
                                          ("dicow_logits_rules_args_t", "LogitsRulesArgs"),
                                          ("dicow_softlabel_ce_args_t", "SoftlabelCeArgs"),
                                          ("dicow_ctc_loss_args_t", "CtcLossArgs"),
                                          ("dicow_decode_linear_args_t", "DecodeLinearArgs"),
                                          ("dicow_ctc_joint_args_t", "CtcJointArgs"),
                                          ("dicow_beam_step_args_t", "BeamStepArgs"),
                                          ("dicow_augment_args_t", "AugmentArgs"),
                                          ("dicow_decode_layer_args_t", "DecodeLayerArgs"),
                                          ("dicow_decode_layers_args_t", "DecodeLayersArgs"),
                                          ("dicow_adamw_tensor_args_t", "AdamwTensorArgs")])
def test_struct_mirrors(lib, cname, pyname):
    m = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + r"\s*;", _header_text(), flags=re.S)
    assert m, cname
    fields = []
    for decl in m.group(1).split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = decl.split(",")
        first = names[0].split()[-1]
        fields.append(first.lstrip("*"))
        fields += [n.strip().lstrip("*") for n in names[1:]]
    py = [f[0] for f in getattr(lib, pyname)._fields_]
    assert py == fields, f"{cname}: header {fields} vs ctypes {py}"


def test_every_header_struct_has_a_mirror_of_the_same_size(lib, tmp_path):
    """sizeof() of every args struct as gcc lays it out from include/dicow_b200.h == ctypes.sizeof of its mirror: catches a
    field type / padding drift that the name comparison above cannot see (the library checks struct_size at run time, but
    only on a GPU box)"""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    names = re.findall(r"\}\s*(dicow_\w+_args_t)\s*;", _header_text())
    assert len(names) >= 17
    mirrors = {}
    for attr in dir(lib):
        obj = getattr(lib, attr)
        if isinstance(obj, type) and issubclass(obj, C.Structure) and obj is not C.Structure:
            key = "dicow_" + re.sub(r"(?<!^)(?=[A-Z])", "_", attr[:-4]).lower() + "_args_t"  # GemmSkinnyArgs -> dicow_gemm_skinny_args_t
            mirrors[key] = obj
    assert sorted(mirrors) == sorted(names), (sorted(set(names) ^ set(mirrors)))
    src = tmp_path / "sizes.c"
    body = "".join(f'  printf("{n} %zu\\n", sizeof({n}));\n' for n in names)
    src.write_text(f'#include <stdio.h>\n#include "{HEADER}"\nint main(void) {{\n{body}  return 0;\n}}\n')
    exe = tmp_path / "sizes"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    for line in out.strip().splitlines():
        n, size = line.split()
        assert C.sizeof(mirrors[n]) == int(size), f"{n}: header {size} bytes, ctypes {C.sizeof(mirrors[n])}"


def test_plain_c_client_links_and_fails_loudly_without_a_device(lib, tmp_path):
    """the boundary is a C ABI: a C99 program that only sees include/dicow_b200.h links against libdicow_b200.so, reads the
    ABI version, and -- in this container, without a GPU -- gets a non-zero status from dicow_create instead of a handle"""
    import shutil
    import subprocess
    import torch
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from ts_asr_whisper_b200 import build
    so = build.build()
    src = tmp_path / "client.c"
    src.write_text(f"""#include <stdio.h>
#include "{HEADER}"
int main(void) {{
  dicow_handle_t h = 0;
  int rc = dicow_create(0, &h);
  printf("abi %d create %d handle %d\\n", dicow_abi_version(), rc, h != 0);
  if (rc == 0) dicow_destroy(h);
  return 0;
}}
""")
    exe = tmp_path / "client"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src), so, f"-Wl,-rpath,{os.path.dirname(so)}"], check=True)
    env = dict(os.environ)
    cuda_lib = "/usr/local/cuda/lib64"
    env["LD_LIBRARY_PATH"] = cuda_lib + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True, env=env).stdout.split()
    assert out[0] == "abi" and int(out[1]) >= 1
    if torch.cuda.is_available():
        assert int(out[3]) == 0 and int(out[5]) == 1
    else:
        assert int(out[3]) != 0 and int(out[5]) == 0


def test_no_device_fails_loudly(lib):
    """without an sm_100 GPU the handle cannot be created and ops raise -- there is no CPU fallback"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ts_asr_whisper_b200 import ops
    with pytest.raises(lib.DicowError):
        lib.handle(0)
    with pytest.raises(ops.DicowError):
        ops.cast_bf16(torch.zeros(4))


def test_oracle_is_imported_by_test_infrastructure_only():
    """only tests/, __graft_entry__ (build / smoke) and the CPU-baseline leg of bench.py may import oracle/: the product
    package and the tools never do (a product path routed through the oracle would void every parity claim)"""
    import ast
    import glob

    def oracle_imports(path):
        tree = ast.parse(open(path).read())
        hits = []
        for node in ast.walk(tree):
            if isinstance(node, ast.Import) and any(a.name.split(".")[0] == "oracle" for a in node.names):
                hits.append(node)
            if isinstance(node, ast.ImportFrom) and (node.module or "").split(".")[0] == "oracle" and node.level == 0:
                hits.append(node)
        return tree, hits
    for path in glob.glob(os.path.join(ROOT, "ts-asr-whisper_b200", "**", "*.py"), recursive=True) + \
            glob.glob(os.path.join(ROOT, "tools", "*.py")):
        assert not oracle_imports(path)[1], f"{path} imports oracle/"
    tree, hits = oracle_imports(os.path.join(ROOT, "bench.py"))
    assert hits
    allowed = [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "time_reference"]
    assert allowed
    inside = {id(n) for n in ast.walk(allowed[0])}
    assert all(id(h) in inside for h in hits), "bench.py imports oracle/ outside the CPU-baseline / reference leg"


def test_gpu_side_never_reads_the_reference_tree():
    """/root/reference does not exist on the GPU box: the -m gpu tests, smoke(), bench.py and the product package must not
    open it (docstrings cite reference file:line, but no path under /root/reference appears as a string to open)"""
    import glob
    files = glob.glob(os.path.join(ROOT, "tests", "test_gpu_*.py")) + [os.path.join(ROOT, "bench.py"),
                                                                        os.path.join(ROOT, "__graft_entry__.py")]
    files += glob.glob(os.path.join(ROOT, "ts-asr-whisper_b200", "*.py"))
    import ast
    for path in files:
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):  # drop docstrings (they cite the reference); comments vanish in ast.unparse
            body = getattr(node, "body", None)
            if isinstance(body, list) and body and isinstance(body[0], ast.Expr) and isinstance(body[0].value, ast.Constant) \
                    and isinstance(body[0].value.value, str):
                body[0].value.value = ""
        assert "/root/reference" not in ast.unparse(tree), path
