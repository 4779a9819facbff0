"""Fill the R2_* placeholders of DESIGN.md / README.md from the final bench lines under profiles/ (maintenance helper)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
L = lambda f: json.loads(open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1])
t, i, t5 = L(sys.argv[1]), L(sys.argv[2]), L(sys.argv[3])
cb = t["cudnn_baseline"]
best = max(cb["fp32"]["value"], cb["bf16_autocast_channels_last"]["value"])
rep = {
    "R2_TRAIN_V": f"{t['value']:.0f}", "R2_TRAIN_MS": f"{t['ms_per_step']:.2f}", "R2_TRAIN_E2E": f"{t['e2e']['value']:.0f}",
    "R2_CUDNN_F32": f"{cb['fp32']['value']:.0f}", "R2_CUDNN_BF16": f"{cb['bf16_autocast_channels_last']['value']:.0f}",
    "R2_CUDNN_X": f"{t['value'] / best:.1f}", "R2_CPU": f"{t['cpu_baseline']['value']:.1f}",
    "R2_T512_V": f"{t5['value']:.0f}", "R2_T512_MS": f"{t5['ms_per_step']:.1f}", "R2_T512_ROOF": f"{t5['roofline']['frac']:.2f}",
    "R2_T512_CUDNN": f"{t5['cudnn_baseline']['bf16_autocast_channels_last']['value']:.0f}", "R2_T512_HBM": f"{t5['roofline_hbm']['frac']:.2f}",
    "R2_INF_V": f"{i['value']:.0f}", "R2_INF_MS": f"{i['ms_per_step']:.1f}", "R2_INF_E2E": f"{i['e2e']['value']:.0f}",
    "R2_INF_SYNC": f"{i['e2e_sync']['value']:.0f}",
    "R2_ROOF": f"{t['roofline']['frac']:.2f}", "R2_HBM": f"{t['roofline_hbm']['frac']:.2f}",
    "R2_BN_FRAC": f"{t['roofline_hbm']['batchnorm_only']['frac']:.2f}",
}
for doc in ("DESIGN.md", "README.md"):
    p = os.path.join(ROOT, doc)
    s = open(p).read()
    for k in sorted(rep, key=len, reverse=True):
        s = s.replace(k, rep[k])
    open(p, "w").write(s)
print(rep)
