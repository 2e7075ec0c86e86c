"""Drop-in model package for the reference repo: copy (or symlink) this directory to
``<reference>/models/fastenhancer_b200/`` and set ``model: "fastenhancer_b200"`` in the YAML
(the reference resolves ``models.<model>.model.Model`` in wrappers/ns.py:29-32 and ``.ONNXModel`` in
scripts/export_onnx.py:32-35).  Requires the ``fastenhancer_b200`` package (this repo) on PYTHONPATH."""
from fastenhancer_b200.model import Model, ONNXModel, StreamingModel  # noqa: F401
