"""CPU: the product path (package, C ABI library, example, headers) never imports, includes, links or executes anything
under oracle/ -- the oracle is test infrastructure -- and fails loudly without its CUDA library."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRODUCT_DIRS = ["odr_audioenc_b200", "examples", "include"]


def _product_files():
    for d in PRODUCT_DIRS:
        for base, _, names in os.walk(os.path.join(ROOT, d)):
            for n in names:
                if n.endswith((".py", ".cpp", ".cu", ".h", ".c")):
                    yield os.path.join(base, n)


def test_no_product_source_touches_the_oracle():
    pat = re.compile(r'(#\s*include\s*[<"][^>"]*oracle|import\s+oracle|from\s+oracle|oracle/|mp2o_|libmp2_oracle|_ref/)')
    hits = []
    for path in _product_files():
        for i, line in enumerate(open(path, errors="replace"), 1):
            code = line.split("//")[0] if path.endswith((".cpp", ".cu", ".h", ".c")) else line.split("#")[0]
            if pat.search(code):
                hits.append("%s:%d: %s" % (os.path.relpath(path, ROOT), i, line.strip()))
    assert not hits, "\n".join(hits)


def test_library_links_only_cuda_and_system_libraries():
    lib = os.path.join(ROOT, "odr_audioenc_b200", "libtoolame_b200.so")
    out = subprocess.run(["ldd", lib], capture_output=True, text=True, check=True).stdout
    assert "oracle" not in out and "toolame_ref" not in out, out
    syms = subprocess.run(["nm", "-D", "--undefined-only", lib], capture_output=True, text=True, check=True).stdout
    assert "mp2o_" not in syms


def test_missing_library_is_an_error_not_a_fallback(tmp_path, monkeypatch):
    from odr_audioenc_b200 import binding
    monkeypatch.setattr(binding, "_lib", None)
    monkeypatch.setattr(binding, "lib_path", lambda: str(tmp_path / "libtoolame_b200.so"))
    try:
        binding.lib()
    except binding.TlbError as e:
        assert "missing" in str(e)
    else:
        raise AssertionError("a missing CUDA library must raise")
