"""Per-layer device time under stage-skip / knob environment settings (timing experiments; results with any
*_DBG bit set are wrong by design).  python tools/stage_sweep.py rp | up6"""
import ctypes as C, json, os, subprocess, sys
SETS = {
    "rp": [("base", {}), ("no_epilogue", {"SRT_RP_DBG": "1"}), ("no_mma", {"SRT_RP_DBG": "2"}), ("no_patch_tma", {"SRT_RP_DBG": "4"}),
           ("no_weight_tma", {"SRT_RP_DBG": "8"}), ("no_mma_no_epi", {"SRT_RP_DBG": "3"}), ("only_mma", {"SRT_RP_DBG": "13"}),
           ("only_epi", {"SRT_RP_DBG": "14"})],
    "up6": [("base", {}), ("lane0_poll", {"SRT_UP6_DBG": "16"}), ("all_off", {"SRT_UP6_DBG": "15"}), ("no_mma", {"SRT_UP6_DBG": "2"}),
            ("no_gather", {"SRT_UP6_DBG": "1"}), ("simt", {"SRT_UP6_TC": "0"})],
    "up6b": [("base", {}), ("stages4", {"SRT_UP6_STAGES": "4"}), ("stages3", {"SRT_UP6_STAGES": "3"}), ("acc8", {"SRT_UP6_ACC": "8"}),
             ("pf24", {"SRT_UP6_PREFETCH": "24"}), ("pf6", {"SRT_UP6_PREFETCH": "6"}), ("no_gather", {"SRT_UP6_DBG": "1"}), ("all_off", {"SRT_UP6_DBG": "15"})],
    "up6c": [("base", {}), ("rna_split", {"SRT_UP6_DBG": "32"}), ("no_gather", {"SRT_UP6_DBG": "1"}), ("no_split", {"SRT_UP6_DBG": "8"})],
    "epw": [("base", {}), ("down1_8warps", {"SRT_RP_DBG": "64"}), ("rp_16warps", {"SRT_RP_DBG": "128"}),
            ("only_epi", {"SRT_RP_DBG": "14"}), ("only_epi_d1_8w", {"SRT_RP_DBG": "78"}), ("only_epi_rp16", {"SRT_RP_DBG": "142"})],
    "mt": [("mt_off", {"SRT_TC_MT": "0"}), ("mt_256", {"SRT_TC_MT": "256"}), ("mt_128", {"SRT_TC_MT": "128"}), ("mt_64", {"SRT_TC_MT": "64"})],
    "istft": [("auto", {}), ("hops16", {"SRT_ISTFT_HOPS": "16"}), ("hops28", {"SRT_ISTFT_HOPS": "28"}), ("hops40", {"SRT_ISTFT_HOPS": "40"}),
              ("hops55", {"SRT_ISTFT_HOPS": "55"})],
}
if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import spleeterrt_b200 as srt
    from spleeterrt_b200 import workload as W
    T, F, N, ns = 512, 1024, 441000, 32
    nets, _ = W.four_stem_nets(); S = len(nets)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    sep = srt.Separator(nets, T, F, max_images=ns, max_batch_images=ns, device=0, cuda_stream=stream.cuda_stream)
    din = torch.randn((ns, 2, N), device="cuda") * 0.1
    dout = torch.empty((ns, S, 2, N), device="cuda")
    n_arr = (C.c_size_t * ns)(*([N] * ns))
    pl = (C.c_void_p * ns)(*[din[i, 0].data_ptr() for i in range(ns)]); pr = (C.c_void_p * ns)(*[din[i, 1].data_ptr() for i in range(ns)])
    po = (C.c_void_p * (ns * S * 2))(*[dout[i, s, c].data_ptr() for i in range(ns) for s in range(S) for c in range(2)])
    for _ in range(3): sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
    torch.cuda.synchronize(); sep.set_timing(True)
    for _ in range(5): sep.separate_raw(pl, pr, n_arr, ns, None, po, device=True)
    torch.cuda.synchronize()
    names = list(W.LAYER_FLOP_PER_PIXEL) + ["down1", "up6", "up7", "stft", "istft"]
    print("RESULT", sys.argv[2], json.dumps({k: round(sep.timing(k) / 5, 3) for k in names}))
else:
    for tag, env in SETS[sys.argv[1]]:
        e = dict(os.environ); e.update(env)
        out = subprocess.run([sys.executable, __file__, "--child", tag], env=e, capture_output=True, text=True, timeout=120)
        print(([l for l in out.stdout.splitlines() if l.startswith("RESULT")] or [out.stderr[-300:]])[0], flush=True)
