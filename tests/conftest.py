import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference mounted (build container only)")


def pytest_collection_modifyitems(config, items):
    have_ref = os.path.isdir("/root/reference/functions")
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="/root/reference not mounted"))
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))


K1_CASES = ["classic18", "cat49_catalog_example", "cat49_brown2018_test", "freeebv_xray49",
            "xray_gamma49", "radio1_sn12_49", "radio2_sn12_49", "radio4_sn12_49", "agnoff18", "radio_xray49",
            "example_thb21_skirtor1p_49", "restrictive18", "midir_uv18", "midir_uv_sn12_49", "xray_midir49"]
K2_CASES = ["default18", "example49", "families18"]
SED_CASES = ["classic18", "freeebv_xray49", "xray_gamma49", "radio_r06_49", "radio1_kd18_49", "radio2_sn12_49"]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
