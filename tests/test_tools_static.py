"""CPU-only checks of the static-evidence tools (tools/ptxas_summary.py)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "tools", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


LOG = """ptxas info    : Compiling entry function '_ZN4mhdf10k_divcleanIfEEvNS_8SpecGeomIT_EEPNS_3CxTIS2_E4typeE' for 'sm_100a'
ptxas info    : Function properties for _ZN4mhdf10k_divcleanIfEEvNS_8SpecGeomIT_EEPNS_3CxTIS2_E4typeE
    24 bytes stack frame, 20 bytes spill stores, 16 bytes spill loads
ptxas info    : Used 28 registers, used 1 barriers, 1280 bytes smem
"""


def test_ptxas_summary_parses_registers_spills_and_smem(tmp_path):
    T = _load("ptxas_summary")
    p = tmp_path / "ptxas.log"
    p.write_text(LOG)
    rows = T.parse(str(p))
    assert list(rows) == ["k_divclean<float>"]
    r = rows["k_divclean<float>"]
    assert (r["regs"], r["stack"], r["st"], r["ld"], r["smem"]) == (28, 24, 20, 16, 1280)


def test_ptxas_summary_names_match_the_committed_summary_spelling():
    T = _load("ptxas_summary")
    assert T.short("void mhdf::k_pass<float, 256, 16, 8, -1, false, 2, 4>(mhdf::PassArgs<float>)") == "k_pass<float, 256, 16, 8, -1, 0, 2, 4>"
    assert T.short("void mhdf::k_xfused<float, 256, 8, 4, (mhdf::Phys)1, true, false>(int)") == "k_xfused<float, 256, 8, 4, 1, 1, 0>"
    old = T.read_summary(os.path.join(ROOT, "profiles", "r01_ptxas_summary_final.txt"))
    assert old["k_xfused<float, 256, 8, 4, 1, 1, 0>"][0] == 255 and len(old) > 300
