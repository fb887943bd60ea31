"""Development probe: where does the time between `ms_per_step` and the kernel's own duration go?
Runs the device-resident step with and without the short last frame (which adds a second launch and an append kernel)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import zigflac_b200 as zf
bits, rate = 24, 96000
for n in (57600000, 57600000 // 4096 * 4096):
    pcm = zf.synth_pcm(n, rate, bits)
    d_pcm = torch.from_numpy(pcm).cuda()
    frames = (n + 4095) // 4096
    enc = zf.Encoder(zf.Config.default(2, bits), rate, max_frames_per_batch=frames)
    cap = enc.max_batch_bytes(frames)
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    d_sizes = torch.zeros(frames, dtype=torch.int32, device="cuda")
    d_total = torch.zeros(1, dtype=torch.int64, device="cuda")
    st = torch.cuda.Stream()
    def step():
        enc.encode_device(d_pcm.data_ptr(), n, 0, d_out.data_ptr(), cap, d_sizes.data_ptr(), d_total.data_ptr(), st.cuda_stream)
    for _ in range(5): step()
    torch.cuda.synchronize(); enc.kernel_times()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(50): step()
    e1.record(st); torch.cuda.synchronize()
    kt = enc.kernel_times()
    print("samples", n, "tail", n % 4096, "ms/step %.4f kernel %.4f gap %.1f us" % (e0.elapsed_time(e1) / 50, sum(kt) / len(kt), (e0.elapsed_time(e1) / 50 - sum(kt) / len(kt)) * 1000))
    enc.close()
