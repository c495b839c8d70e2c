import os, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, os.getcwd())
import numpy as np, torch, cv2
from oracle import ref_harness as R
from dad_3dheads_b200.encoder_weights import synthetic_state_dict
sd = synthetic_state_dict(0)
img = cv2.cvtColor(cv2.imread("tests/golden/demo_head_1.jpeg"), cv2.COLOR_BGR2RGB)
ref = R.predictor(sd)
cache = {}
x = ref.preprocess(img, cache)
with torch.no_grad():
    traced = ref.model(x)["OUTPUT_3DMM_PARAMS"]
    eager = R.flame_regression(sd)(x)["OUTPUT_3DMM_PARAMS"]
    eager64 = R.flame_regression(sd, dtype=torch.float64)(x.double())["OUTPUT_3DMM_PARAMS"]
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
print("threads", torch.get_num_threads(), "traced vs eager64", rel(traced, eager64), "eager32 vs eager64", rel(eager, eager64))
z = np.load("tests/golden/reference_predictor.npz")
res = ref(img.copy())
print("live traced predictor vs committed fixture: params", rel(res["3dmm_params"], torch.from_numpy(z["params_3dmm"])))
if torch.cuda.is_available():
    from dad_3dheads_b200.predictor import FaceMeshPredictor
    for prec in ("fp32", "fp16x2"):
        pred = FaceMeshPredictor.dad_3dnet(state_dict=sd, precision=prec)
        got = pred.model(x.cuda())["OUTPUT_3DMM_PARAMS"].cpu()
        print(prec, "gpu vs eager64", rel(got, eager64), "gpu vs traced", rel(got, traced))
        out = pred(img.copy())
        print(prec, "call vs fixture params", rel(out["3dmm_params"], torch.from_numpy(z["params_3dmm"])), "verts", rel(out["3d_vertices"], torch.from_numpy(z["vertices_3d"])))
