"""Times the full search at the BASELINE size for a few gallery-range schedules (tuning aid, not the bench)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visiondk_b200.retrieval import FlatIPIndex

nq, ng, dim, k = 10000, 1000000, 512, 100
gen = torch.Generator(device="cuda").manual_seed(5)
g = torch.nn.functional.normalize(torch.randn(ng, dim, device="cuda", generator=gen))
q = torch.nn.functional.normalize(torch.randn(nq, dim, device="cuda", generator=gen))
idx = FlatIPIndex(dim, "cuda")
idx.add(g)
ref = None
for name, ends in [("default 4k,32k,262k", None), ("4k,16k,64k,256k", [4096, 16384, 65536, 262144]),
                   ("8k,64k,512k", [8192, 65536, 524288]), ("4k,32k", [4096, 32768]), ("8k,128k", [8192, 131072]),
                   ("4k,12k,40k,128k,400k", [4096, 12288, 40960, 131072, 409600]), ("16k,256k", [16384, 262144])]:
    idx.stage_ends = ends
    try:
        for _ in range(3):
            s, i = idx.search_device(q, k)
        st = idx.check_status()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            s, i = idx.search_device(q, k)
        e1.record()
        torch.cuda.synchronize()
        same = True if ref is None else bool(torch.equal(i, ref))
        if ref is None:
            ref = i.clone()
        print(json.dumps({"schedule": name, "ms": e0.elapsed_time(e1) / 10, "status": st, "same_ids": same}))
    except Exception as ex:
        print(json.dumps({"schedule": name, "error": str(ex)[:200]}))
