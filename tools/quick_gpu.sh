cd $GRAFT_REPO_ROOT
python - <<'PY'
import torch
t = torch.empty(1 << 31, dtype=torch.uint8, device="cuda")
for name, fn in (("fill (write only)", lambda: t.fill_(3)), ("copy (read+write)", lambda: t[: 1 << 30].copy_(t[1 << 30:]))):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print("%-20s %.1f GB/s" % (name, (2 ** 31) / ms / 1e6))
PY
python tools/validate_quality.py puzzle15 40 fp16x3 2>&1 | tail -10
python tools/validate_quality.py cube3 100 fp16x3 2>&1 | tail -10
