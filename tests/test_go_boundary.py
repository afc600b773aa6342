"""The Go side of the drop-in boundary cannot be compiled here (no Go toolchain in the image), so it is
checked mechanically instead (tools/check_go_boundary.py): every exported identifier of the reference
files INTEGRATION.md excludes under `-tags sdr.cuda` is re-declared by the twins with the same signature,
every C symbol the Go code calls exists in the header and every header function is bound, nothing the
excluded files alone declare is still used, and INTEGRATION.md names only symbols that exist."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import check_go_boundary as G  # noqa: E402

REFERENCE = "/root/reference"


def test_every_c_symbol_is_declared_and_bound():
    assert G.check_b() == []


def test_integration_md_names_only_real_symbols():
    assert G.check_d() == []


def test_twins_redeclare_every_excluded_exported_identifier():
    """Against the committed expectation list (generated from the reference by --write-expected), and
    against the reference itself when it is mounted."""
    with open(os.path.join(ROOT, "tests", "golden", "go_boundary_expected.json")) as fh:
        expect = json.load(fh)
    assert sum(len(v) for v in expect.values()) >= 20
    for name in ("Copy", "CopyBuffer", "ErrConversionNotImplemented", "ConvertBuffer", "CopySamples"):
        assert name in expect["root"]
    for name in ("ShiftBuffer", "DecimateBuffer", "DownsampleBuffer", "ReadBeamform", "Beamform.SetPhaseAngles"):
        assert name in expect["stream"]
    assert G.check_a(expect) == []
    if os.path.isdir(REFERENCE):
        live = G.reference_expectations(REFERENCE)
        assert live == expect, "tests/golden/go_boundary_expected.json is stale: tools/check_go_boundary.py --write-expected"
        assert G.check_a(live) == []


def test_nothing_left_dangling_in_the_reference_packages():
    if not os.path.isdir(REFERENCE):
        import pytest
        pytest.skip("reference tree not mounted")
    assert G.check_c(REFERENCE) == []


def test_signature_parser():
    d = G.declarations("""
package x
func ConvertWriter(
	out sdr.Writer,
	inputFormat sdr.SampleFormat,
) (sdr.Writer, error) {
}
func ShiftBuffer(sampleRate uint) func(rf.Hz, sdr.SamplesC64) {
}
func (b *Beamform) SetPhaseAngles(angles []complex64) error {
}
func Add(readers ...sdr.Reader) (sdr.Reader, error) {
}
func f(a, b int, c string) (n int, err error) {
}
var (
	ErrA = 1
	errB = 2
)
type T struct{}
""")
    assert d["ConvertWriter"] == "func(sdr.Writer,sdr.SampleFormat)(sdr.Writer,error)"
    assert d["ShiftBuffer"] == "func(uint)(func(rf.Hz,sdr.SamplesC64))"
    assert d["Beamform.SetPhaseAngles"] == "func([]complex64)(error)"
    assert d["Add"] == "func(...sdr.Reader)(sdr.Reader,error)"
    assert d["f"] == "func(int,int,string)(int,error)"
    assert d["ErrA"] == "var" and d["errB"] == "var" and d["T"] == "type"
