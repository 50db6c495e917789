"""PCIe copy-rate probe for the host-pointer API design (run on the GPU box): pinned H2D / D2H rates for one
large copy vs. the per-(stream, stem, channel) pieces srt_separate_batch issues, alone, duplex and under compute."""
import json, time, torch

def rate(fn, nbytes, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9

N = 441000
h_out = torch.empty((256, N), dtype=torch.float32).pin_memory()
d_out = torch.empty((256, N), dtype=torch.float32, device="cuda")
h_in = torch.empty((64, N), dtype=torch.float32).pin_memory()
d_in = torch.empty((64, N), dtype=torch.float32, device="cuda")
s_in, s_out, s_c = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
a = torch.randn(8192, 8192, device="cuda"); b = torch.randn(8192, 8192, device="cuda")
res = {}
def d2h_big():
    with torch.cuda.stream(s_out): h_out.copy_(d_out, non_blocking=True)
def d2h_pieces():
    with torch.cuda.stream(s_out):
        for i in range(256): h_out[i].copy_(d_out[i], non_blocking=True)
def h2d_big():
    with torch.cuda.stream(s_in): d_in.copy_(h_in, non_blocking=True)
def duplex():
    h2d_big(); d2h_big()
def compute():
    with torch.cuda.stream(s_c):
        for _ in range(4): torch.mm(a, b)
def d2h_under_compute():
    compute(); d2h_big()
def duplex_under_compute():
    compute(); h2d_big(); d2h_big()
nb = h_out.numel() * 4
res["d2h_one_copy_gbs"] = rate(d2h_big, nb)
res["d2h_256_pieces_gbs"] = rate(d2h_pieces, nb)
res["h2d_one_copy_gbs"] = rate(h2d_big, h_in.numel() * 4)
res["d2h_while_h2d_gbs"] = rate(duplex, nb)
res["d2h_under_compute_gbs"] = rate(d2h_under_compute, nb)
res["d2h_duplex_under_compute_gbs"] = rate(duplex_under_compute, nb)
print(json.dumps(res))
