import sys, os, time, json
sys.path.insert(0, "/root/repo")
import torch, pynvml
from picopose_b200 import matching as M, synth
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
dev = "cuda:0"
B, N, C, H = 64, 162, 1024, 16
g = torch.Generator(device=dev).manual_seed(0)
bank_f = torch.randn(1, N, C, H, H, device=dev, generator=g)
tar = bank_f[0, :B].clone() + 0.5 * torch.randn(B, C, H, H, device=dev, generator=g)
bank = M.TemplateBank.from_features(bank_f)
bidx = torch.zeros(B, dtype=torch.int32, device=dev)
for mask_name, mask in (("disc", synth.disc_mask(B).to(dev)), ("ones", torch.ones(B, 224, 224, device=dev))):
    for iters in (1, 5, 20, 100):
        for _ in range(3):
            M.matching_templates(bank, tar, None, mask, topk=5, bank_index=bidx)
        torch.cuda.synchronize()
        time.sleep(0.2)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
        t0 = time.perf_counter()
        evs[0].record()
        for i in range(iters):
            M.matching_templates(bank, tar, None, mask, topk=5, bank_index=bidx)
            evs[i + 1].record()
        t_enq = time.perf_counter() - t0
        clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
        torch.cuda.synchronize()
        per = [evs[i].elapsed_time(evs[i + 1]) for i in range(iters)]
        print(mask_name, iters, "mean %.3f ms  first %.3f last %.3f  enqueue/call %.3f ms  clk %d MHz  %.0f W" % (
            sum(per) / iters, per[0], per[-1], t_enq / iters * 1e3, clk, pw), flush=True)
