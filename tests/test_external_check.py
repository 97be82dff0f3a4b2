"""tools/export_external_check.py: the package that lets ONE run of the real `fade annotate` elsewhere pin
(or refute) the oracle.  CPU: the package is complete and its comparer accepts the oracle's own
annotation and rejects a corrupted one, naming the switch that explains a U-variant.  GPU: the C++
driver's output on the exported inputs passes the comparer (fade-b200 standing in for fade)."""
import importlib.util
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "fade_b200", "bin", "fade-b200")


def _tool():
    spec = importlib.util.spec_from_file_location("export_external_check", os.path.join(ROOT, "tools", "export_external_check.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _sam_with_tags(src_sam, tsv, dst_sam, reverse=True):
    """the input SAM + the tags of `tsv`, as `fade annotate` writes them (rs always; am/as/ar/ab for artifacts)"""
    exp = {}
    with open(tsv) as f:
        next(f)
        for line in f:
            q, flag, rs, am, as_, ar, ab = line.rstrip("\n").split("\t")
            exp[(q, flag)] = (rs, am, as_, ar, ab)
    header, body = [], []
    for line in open(src_sam):
        if line.startswith("@"):
            header.append(line)
            continue
        fl = line.rstrip("\n").split("\t")
        rs, am, as_, ar, ab = exp[(fl[0], fl[1])]
        fl.append(f"rs:i:{rs}")
        if int(rs) & 6:
            fl += [f"am:Z:{am}", f"as:Z:{as_}", f"ar:Z:{ar}", f"ab:Z:{ab}"]
        body.append("\t".join(fl) + "\n")
    if reverse:
        body.reverse()      # fade's output order is unspecified
    with open(dst_sam, "w") as f:
        f.writelines(header + body)


def test_package_and_comparer(tmp_path):
    tool = _tool()
    out = tmp_path / "pkg"
    info = tool.export(str(out), 600, variants=True)
    assert info["artifact_records"] > 20
    for f in ("ref.fa", "reads.sam", "expected_tags.tsv", "compare_external.py", "README.txt"):
        assert (out / f).stat().st_size > 0
    cmp_py = str(out / "compare_external.py")
    good = tmp_path / "fade_out.sam"
    _sam_with_tags(out / "reads.sam", out / "expected_tags.tsv", good)
    p = subprocess.run([sys.executable, cmp_py, str(good)], capture_output=True, text=True)
    assert p.returncode == 0 and "PARITY OK" in p.stdout, p.stdout + p.stderr
    # a run that behaves like one of the documented variants is recognised as such
    alt = tmp_path / "alt.sam"
    _sam_with_tags(out / "reads.sam", out / "expected_tags.U1_no_softclip_pad.tsv", alt)
    p = subprocess.run([sys.executable, cmp_py, str(alt)], capture_output=True, text=True)
    assert p.returncode == 1 and "U1_no_softclip_pad.tsv: 0 records differ  <-- this switch explains the run" in p.stdout, p.stdout
    # and plain corruption is not explained by anything
    bad = tmp_path / "bad.sam"
    txt = open(good).read().replace("rs:i:3", "rs:i:1", 1)
    open(bad, "w").write(txt)
    p = subprocess.run([sys.executable, cmp_py, str(bad)], capture_output=True, text=True)
    assert p.returncode == 1 and "explains the run" not in p.stdout


@pytest.mark.gpu
def test_cli_output_passes_the_external_comparer(tmp_path):
    tool = _tool()
    out = tmp_path / "pkg"
    tool.export(str(out), 3000, variants=False)
    res = tmp_path / "fade_out.sam"
    with open(res, "w") as fo:
        p = subprocess.run([BIN, "annotate", str(out / "reads.sam"), str(out / "ref.fa")], stdout=fo, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    p = subprocess.run([sys.executable, str(out / "compare_external.py"), str(res)], capture_output=True, text=True)
    assert p.returncode == 0 and "PARITY OK" in p.stdout, p.stdout
