"""The C-ABI library loads, exports every symbol include/mjpl_b200.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls here."""

import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from mjpl_b200 import _abi, models

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "mjpl_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mjb_[a-z_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    L = _abi.lib()
    names = _declared()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/mjpl_b200.h but not exported"
    assert sorted(_abi.EXPORTS) == names


def test_desc_layout_matches_header_field_order():
    text = (ROOT / "include" / "mjpl_b200.h").read_text()
    body = text[text.index("typedef struct mjb_model_desc {"): text.index("} mjb_model_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        names = re.sub(r"^(const\s+)?(int32_t|int64_t|double)\s*", "", decl)
        fields += [n.strip().lstrip("*") for n in names.split(",")]
    assert fields == [f[0] for f in _abi.ModelDesc._fields_]


def test_bad_arguments_are_rejected_before_any_cuda_work():
    L = _abi.lib()
    out = C.c_void_p()
    assert L.mjb_model_create(None, C.byref(out)) == _abi.MJB_ERR_ARG
    # unsupported joint type is a model error even without a device
    from mjpl_b200 import mjcf

    from . import toy_models as toys

    m = mjcf.from_xml_string(toys.JOINT_ZOO)
    desc, keep = _abi.make_desc(m, [])
    assert L.mjb_model_create(C.byref(desc), C.byref(out)) == _abi.MJB_ERR_MODEL
    assert b"hinge and slide" in L.mjb_last_error()


def test_no_gpu_means_loud_failure():
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    L = _abi.lib()
    assert L.mjb_device_count() == 0
    m = models.load("two_dof_ball")
    desc, keep = _abi.make_desc(m, [])
    out = C.c_void_p()
    rc = L.mjb_model_create(C.byref(desc), C.byref(out))
    assert rc == _abi.MJB_ERR_CUDA and b"no CPU fallback" in L.mjb_last_error()
    import mjpl_b200 as mj

    with pytest.raises(mj.EngineUnavailable):
        mj.CollisionConstraint(m)
    with pytest.raises(mj.EngineUnavailable):
        mj.JointLimitConstraint(m)
    with pytest.raises(mj.EngineUnavailable):
        mj.get_engine(m).valid_configs(np.zeros((1, 2)))


def test_product_never_imports_the_oracle():
    for p in (ROOT / "mjpl_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            t = p.read_text()
            assert "import oracle" not in t and "from oracle" not in t and "oracle.h" not in t, p
