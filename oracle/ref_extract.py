"""Run the reference's OWN function bodies on torch-CPU — golden-vector generation only.

TEST INFRASTRUCTURE. Used by tests/golden/make_golden.py in the build container, where
/root/reference exists. Nothing here is imported by the product, by ``-m gpu`` tests, smoke() or
bench.py (the GPU box has no /root/reference).

The reference package cannot be imported (mmcv / mmdet / spconv / torch_scatter are absent and
mmdet3d/__init__.py asserts an mmcv version), so:

* detector *methods* (voxelize_points, point_to_cam, sample_points_triplane, roi) are cut out of
  their class with ``ast`` and compiled as free functions taking a duck-typed ``self``;
* the ``PointTriplaneProjector`` module file is imported as-is with three stub modules injected into
  ``sys.modules``: ``mmdet.models.builder`` (registry decorator), ``torch_scatter`` and
  ``spconv.pytorch`` (the restatements in oracle/triplane_oracle.py).

No reference source text is copied into this repository; it is read from /root/reference at run time.
"""
from __future__ import annotations

import ast
import importlib.util
import math
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REFERENCE_ROOT = os.environ.get("TP_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mmdet3d"))


def load_method(relpath: str, class_name: str, method_name: str):
    """Return ``class_name.method_name`` of a reference file as a plain function(self, ...)."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path, "r") as fh:
        tree = ast.parse(fh.read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == method_name:
                    mod = ast.Module(body=[item], type_ignores=[])
                    ns = {"torch": torch, "F": F, "np": np, "math": math}
                    exec(compile(mod, path, "exec"), ns)
                    return ns[method_name]
    raise LookupError(f"{class_name}.{method_name} not found in {path}")


def load_statements(relpath: str, class_name: str, method_name: str, contains: str, before: int = 0):
    """Compile a run of statements from INSIDE a reference method as a function(namespace) -> namespace: the first
    statement (searched depth-first) whose source contains `contains`, plus the `before` sibling statements in front of
    it. For code the reference writes inline in forward() (e.g. the camera-pixel scatter, triplane.py:379-390)."""
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path, "r") as fh:
        src = fh.read()
    tree = ast.parse(src, filename=path)

    def find(body):
        for k, st in enumerate(body):
            seg = ast.get_source_segment(src, st) or ""
            if contains in seg:
                for field in ("body", "orelse"):
                    sub = getattr(st, field, None)
                    if isinstance(sub, list) and sub and isinstance(st, (ast.If, ast.With)):
                        hit = find(sub)
                        if hit is not None:
                            return hit
                return body[max(0, k - before):k + 1]
        return None

    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            for item in node.body:
                if isinstance(item, ast.FunctionDef) and item.name == method_name:
                    stmts = find(item.body)
                    if stmts is None:
                        break
                    code = compile(ast.Module(body=list(stmts), type_ignores=[]), path, "exec")

                    def run(ns):
                        ns = dict(ns)
                        ns.update({"torch": torch, "F": F, "np": np, "math": math})
                        exec(code, ns)
                        return ns

                    run.lines = (stmts[0].lineno, stmts[-1].end_lineno)
                    return run
    raise LookupError(f"no statement containing {contains!r} in {class_name}.{method_name} of {path}")


class _Registry:
    def register_module(self, *a, **k):
        return lambda cls: cls


def _install_stub_modules():
    from oracle import triplane_oracle as O

    mmdet = types.ModuleType("mmdet")
    mmdet_models = types.ModuleType("mmdet.models")
    mmdet_builder = types.ModuleType("mmdet.models.builder")
    mmdet_builder.BACKBONES = _Registry()
    mmdet.models = mmdet_models
    mmdet_models.builder = mmdet_builder

    ts = types.ModuleType("torch_scatter")

    def scatter_max(src, index, dim=0):
        assert dim == 0
        n = int(index.max()) + 1 if index.numel() else 0
        return O.scatter_max(src, index, n), None

    ts.scatter_max = scatter_max

    spconv = types.ModuleType("spconv")
    spconv_pt = types.ModuleType("spconv.pytorch")

    class SparseConvTensor:
        def __init__(self, features, indices, spatial_shape, batch_size):
            self.features, self.indices = features, indices
            self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], int(batch_size)

    class _Pooled:
        def __init__(self, dense):
            self._dense = dense

        def dense(self):
            return self._dense

    class SparseMaxPool3d(torch.nn.Module):
        clamp_zero = False

        def __init__(self, kernel_size, stride=None, padding=0):
            super().__init__()
            assert list(stride) == list(kernel_size) and padding == 0
            self.kernel_size = list(kernel_size)

        def forward(self, x):
            return _Pooled(O.sparse_max_pool_dense(x.features, x.indices, x.spatial_shape, self.kernel_size,
                                                   x.batch_size, clamp_zero=self.clamp_zero))

    spconv_pt.SparseConvTensor = SparseConvTensor
    spconv_pt.SparseMaxPool3d = SparseMaxPool3d
    spconv.pytorch = spconv_pt
    stubs = {"mmdet": mmdet, "mmdet.models": mmdet_models, "mmdet.models.builder": mmdet_builder,
             "torch_scatter": ts, "spconv": spconv, "spconv.pytorch": spconv_pt}
    saved = {k: sys.modules.get(k) for k in stubs}
    sys.modules.update(stubs)
    return saved


def load_projector_class():
    """The reference's PointTriplaneProjector class, third-party ops stubbed (see module docstring)."""
    saved = _install_stub_modules()
    try:
        path = os.path.join(REFERENCE_ROOT, "mmdet3d/models/backbones/point_triplane_projector.py")
        spec = importlib.util.spec_from_file_location("_ref_point_triplane_projector", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod.PointTriplaneProjector


class cpu_randperm:
    """The reference's forward calls torch.randperm(n, device=points[0].get_device()); on CPU that is
    device=-1, which torch rejects. The shuffle does not change the forward result (max / unique are
    permutation invariant), so route it to the CPU generator for the duration of the call."""

    def __enter__(self):
        self._orig = torch.randperm

        def randperm(n, *a, device=None, **k):
            return self._orig(n, *a, **k)

        torch.randperm = randperm
        return self

    def __exit__(self, *exc):
        torch.randperm = self._orig
        return False
