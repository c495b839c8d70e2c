"""-m "not gpu": the reference's module paths resolve through compat/ (what demo.py needs to import)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_module_paths_resolve():
    code = (
        "import sys; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from predictor import FaceMeshPredictor\n"
        "from model_training.head_mesh import HeadMesh\n"
        "from model_training.model.flame import FLAMELayer, FlameParams, FLAME_CONSTS, calculate_rpy\n"
        "from model_training.model.utils import rot_mat_from_6dof, calculate_paddings, to_device, unravel_index\n"
        "from model_training.data.config import OUTPUT_3DMM_PARAMS, OUTPUT_2D_LANDMARKS, OUTPUT_LANDMARKS_HEATMAP\n"
        "from utils import load_yaml, get_relative_path\n"
        "import torch\n"
        "p = torch.zeros(1, 413)\n"
        "p[0, 403:409] = torch.tensor([1., 0, 0, 0, 1., 0])\n"
        "fp = FlameParams.from_3dmm(p, FLAME_CONSTS)\n"
        "r = calculate_rpy(fp)\n"
        "assert abs(r.yaw) < 1e-4 and calculate_paddings(256, 206) == [0, 0, 25, 25]\n"
        "assert FLAME_CONSTS['shape'] == 300 and hasattr(FaceMeshPredictor, 'dad_3dnet')\n"
        "print('ok')\n" % (ROOT, os.path.join(ROOT, "compat")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
