"""Generator for tests/golden/beta_vs_time.json (run once in the build container).

Copies the two numeric arrays of the reference's only golden artefact,
/root/reference/test_scripts/beta_vs_time.mat (written by
test_scripts/stanley_controller_ellipse.py:1066-1069 with CBF_TYPE = 4), into a JSON fixture
with full repr precision, because /root/reference does not exist on the GPU box.
"""
import json
import os

from scipy.io import loadmat

src = "/root/reference/test_scripts/beta_vs_time.mat"
m = loadmat(src)
out = {
    "source": "test_scripts/beta_vs_time.mat",
    "header": m["__header__"].decode(),
    "t_arr": [float(v) for v in m["t_arr"].ravel()],
    "beta_deg": [float(v) for v in m["beta_deg"].ravel()],
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "beta_vs_time.json")
with open(dst, "w") as f:
    json.dump(out, f)
print("wrote", dst, len(out["t_arr"]))
