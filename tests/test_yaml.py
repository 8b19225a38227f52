"""-m "not gpu": the YAML report (main.cpp:214-305 keys through YAML_Doc / YAML_Element) is text-identical to the
one the reference's own classes produce for the same numbers."""
import os

import numpy as np
import pytest


@pytest.mark.parametrize("ranks,omp", [(0, 0), (0, 8), (4, 0)])
def test_yaml_text_identical_to_reference(H, refwrap, tmp_path, ranks, omp):
    if not refwrap.available("serial"):
        pytest.skip("real reference not built")
    times = np.array([12.5, 0.75, 1.25, 10.0, 0.125, 0.25, 0.0625])
    t4 = [0.1, 0.15, 0.125]
    args = (256, 256, 256, 149, 2.2419957139761042e-18, times, 16777216.0 * max(ranks, 1), 27 * 16777216.0 * max(ranks, 1))
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        mine = H.yaml_report(*args, ranks=ranks, omp_threads=omp, t4stats=t4)
    finally:
        os.chdir(old)
    ref = refwrap.yaml_report(*args, ranks=ranks, omp_threads=omp, t4stats=t4, cwd=str(tmp_path))
    assert mine == ref
    assert "Total   : 12.5" in mine and "Number of iterations: 149" in mine
    # both wrote ./hpccg-1.0_<timestamp>.yaml (YAML_Doc.cpp:49-70)
    files = [f for f in os.listdir(tmp_path) if f.startswith("hpccg-1.0_") and f.endswith(".yaml")]
    assert files


def test_yaml_flop_accounting(H, tmp_path):
    """out.txt:29-32: 10^3, 149 iterations -> 9.536e+06 / 596000 / 894000 / 8.046e+06."""
    old = os.getcwd()
    os.chdir(tmp_path)
    try:
        text = H.yaml_report(10, 10, 10, 149, 1e-30, np.ones(7), 1000.0, 27000.0)
    finally:
        os.chdir(old)
    assert "Total   : 9.536e+06" in text and "DDOT    : 596000" in text
    assert "WAXPBY  : 894000" in text and "SPARSEMV: 8.046e+06" in text
