"""The stand-alone driver on the GPU: the reference's case file in, its *.nso file out, every solve through the C ABI.
The rows are compared with the ones the same driver writes when the CPU oracle is its solver."""
import io
import os
import subprocess
import sys
import numpy as np
import pytest

from multifebe_b200.host import shape
from multifebe_b200.host.casefile import CaseFile
from multifebe_b200.host.export import read_nso
from multifebe_b200 import driver
from test_casefile_driver import SOLID_DAT, FLUID_DAT, _write_case, _run_with_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _compare(nso_gpu, nso_cpu, harmonic, first_value_col=12):
    a, b = read_nso(nso_gpu), read_nso(nso_cpu)
    assert a.shape == b.shape and np.array_equal(a[:, :first_value_col], b[:, :first_value_col])
    va, vb = a[:, first_value_col:], b[:, first_value_col:]
    if harmonic:                                   # second half of the value columns = incident field (zero)
        nt = va.shape[1] // 2
        assert not va[:, nt:].any() and not vb[:, nt:].any()
        va, vb = va[:, :nt], vb[:, :nt]
    # primary (u, p) and secondary (t, Un) variables live on different scales: each family against its own maximum (1e-8, BASELINE.json)
    h = va.shape[1] // 2
    for sl in (slice(0, h), slice(h, 2 * h)):
        assert np.abs(va[:, sl] - vb[:, sl]).max() <= 1e-8 * np.abs(vb[:, sl]).max()


def test_static_case(tmp_path):
    text = (SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")).replace("eng_double", "sci_double")
    path = _write_case(tmp_path, text, et=shape.QUAD9, m=2)
    nso_cpu, _ = _run_with_oracle(path, output=path + ".cpu")
    nso = driver.run(path, log=io.StringIO())
    _compare(nso, nso_cpu, False)


def test_harmonic_fluid_and_solid_cases(tmp_path):
    d1 = tmp_path / "fluid"; d1.mkdir()
    path = _write_case(d1, FLUID_DAT + "\n[export]\nreal_format = sci_double\n", et=shape.TRI6, m=2)
    nso_cpu, _ = _run_with_oracle(path, output=path + ".cpu")
    _compare(driver.run(path, log=io.StringIO()), nso_cpu, True)
    d2 = tmp_path / "solid"; d2.mkdir()
    freq = "\n[frequencies]\nrad/s\nlin\n3\n0.5\n6.0\n"
    text = (SOLID_DAT % dict(analysis="harmonic", freq=freq, z="(0.,0.)", one="(1.,0.)")).replace("eng_double", "sci_double")
    path = _write_case(d2, text, et=shape.TRI3, m=3)
    nso_cpu, _ = _run_with_oracle(path, output=path + ".cpu")
    _compare(driver.run(path, log=io.StringIO()), nso_cpu, True)


def test_command_line(tmp_path):
    """python -m multifebe_b200 -i case.dat -o out (options of src/process_command_line_options.f90)."""
    path = _write_case(tmp_path, FLUID_DAT, et=shape.QUAD4, m=3)
    out = str(tmp_path / "result")
    r = subprocess.run([sys.executable, "-m", "multifebe_b200", "-i", path, "-o", out, "-b", "2"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rows = read_nso(out + ".nso")
    md = CaseFile(path).build_model()
    assert rows.shape == (2 * md.n_node, 20) and "frequency 2 / 2 done" in r.stdout
    r = subprocess.run([sys.executable, "-m", "multifebe_b200", "-i", str(tmp_path / "missing.dat")], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 2 and "Input file does not exist" in r.stdout


def test_static_case_with_internal_points(tmp_path):
    text = (SOLID_DAT % dict(analysis="static", freq="", z="0.", one="1.")).replace("eng_double", "sci_double")
    text += "\n[internal points]\n2\n1 1 0.5 0.5 0.5\n2 1 0.2 0.7 0.4\n"
    path = _write_case(tmp_path, text, et=shape.QUAD4, m=3)
    nso_cpu, _ = _run_with_oracle(path, output=path + ".cpu")
    nso = driver.run(path, log=io.StringIO())
    la = [s for s in open(nso) if s.strip() and not s.startswith("#")]
    lb = [s for s in open(nso_cpu) if s.strip() and not s.startswith("#")]
    assert len(la) == len(lb)
    a = np.array([[float(t) for t in s.split()] for s in la[-2:]]); b = np.array([[float(t) for t in s.split()] for s in lb[-2:]])
    assert np.array_equal(a[:, :12], b[:, :12])
    assert np.abs(a[:, 12:15] - b[:, 12:15]).max() <= 1e-8 * np.abs(b[:, 12:15]).max() and np.abs(a[:, 15:] - b[:, 15:]).max() <= 1e-8 * np.abs(b[:, 15:]).max()


def test_case_with_symmetry_planes(tmp_path):
    """[symmetry planes] through the stand-alone driver: the quarter column, harmonic (antisymmetric variant too) and static, against the oracle run."""
    from test_casefile_driver import SYM_DAT, _write_quarter
    freq = "\n[frequencies]\nrad/s\nlin\n2\n0.5\n6.0\n"
    for k, (analysis, planes) in enumerate([("harmonic", "plane_n2: symmetry\nplane_n3: symmetry"), ("harmonic", "plane_n2: antisymmetry\nplane_n3: symmetry"),
                                            ("static", "plane_n2: symmetry\nplane_n3: symmetry")]):
        d = tmp_path / ("c%d" % k); d.mkdir()
        harm = analysis == "harmonic"
        text = (SYM_DAT % dict(analysis=analysis, freq=freq if harm else "", z="(0.,0.)" if harm else "0.", one="(1.,0.)" if harm else "1.", planes=planes)).replace("eng_double", "sci_double")
        path = _write_quarter(d, text, et=shape.TRI6 if k else shape.QUAD9, m=2)
        nso_cpu, _ = _run_with_oracle(path, output=path + ".cpu")
        _compare(driver.run(path, log=io.StringIO()), nso_cpu, harm)


@pytest.mark.parametrize("material,kind,space", [("fluid rho 1.2 c 1.5", "fluid p", "full-space"), ("elastic_solid rho 2. mu 1.5 nu 0.3 xi 0.01", "elastic sh", "half-space 3 5. 1")])
def test_case_with_incident_waves(tmp_path, material, kind, space):
    """[incident waves] through the stand-alone driver on the device (tests/test_incident_waves.py's transparent inclusion: two coupled regions, the field
    in the outer one): the per-frequency arrays reach the H problem of the region, the result file carries the total and the incident field."""
    from test_incident_waves import _transmission_case, _CoupledOracleSolver
    path = _transmission_case(tmp_path, material=material, kind=kind, space=space, varphi="30.", theta="60.")
    case = CaseFile(path)
    nso_cpu = driver.run(path, output=path + ".cpu", solver=_CoupledOracleSolver(case.build_model()), log=io.StringIO())
    nso_gpu = driver.run(path, log=io.StringIO())
    a, b = read_nso(nso_gpu), read_nso(nso_cpu)
    assert a.shape == b.shape and np.array_equal(a[:, :12], b[:, :12])
    nv = (a.shape[1] - 12) // 2
    assert np.array_equal(a[:, 12 + nv:], b[:, 12 + nv:]) and a[:, 12 + nv:].any()           # the incident columns: host arithmetic on both sides
    h = nv // 2
    for sl in (slice(12, 12 + h), slice(12 + h, 12 + nv)):
        assert np.abs(a[:, sl] - b[:, sl]).max() <= 1e-8 * np.abs(b[:, sl]).max()


def test_case_with_bem_formulation_section(tmp_path):
    """[bem formulation over boundaries] on the device: MCA points at every node of two boundaries (sbie_mca, default and given delta), rim displacement 0.01
    on a third -- harmonic and static against the oracle run of the same case."""
    extra = "\n[bem formulation over boundaries]\nboundary 1: sbie_boundary_mca 0.01\nboundary 2: sbie_mca 0.\nboundary 3: sbie_mca 0.3\n"
    freq = "\n[frequencies]\nrad/s\nlist\n2\n0.7\n3.1\n"
    for k, analysis in enumerate(("harmonic", "static")):
        d = tmp_path / ("f%d" % k); d.mkdir()
        harm = analysis == "harmonic"
        text = (SOLID_DAT % dict(analysis=analysis, freq=freq if harm else "", z="(0.,0.)" if harm else "0.", one="(1.,0.)" if harm else "1.")).replace("eng_double", "sci_double") + extra
        path = _write_case(d, text, et=shape.QUAD9 if harm else shape.TRI6, m=2)
        nso_cpu, case = _run_with_oracle(path, output=path + ".cpu")
        assert case.formulation[2] == ("sbie_mca", 0.0)
        _compare(driver.run(path, log=io.StringIO()), nso_cpu, harm)
