"""Turns the artefacts of tools/gpu_full_run.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/."""
import csv, json, subprocess, sys, shutil
tag, out = sys.argv[1], sys.argv[2]
for f in ['bench', 'bench_c1', 'bench_c3', 'bench_ref']:
    d = json.loads(open(f'gpurun_out/{tag}_{f}.json').read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('kernel_ms'), d['e2e']['value'], d.get('gpu_launches'))
    shutil.copy(f'gpurun_out/{tag}_{f}.json', f'profiles/{out}_{f}.json')
shutil.copy(f'gpurun_out/{tag}_launches.csv', f'profiles/{out}_ncu_launches.csv')
WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__icc_request_hit_rate.pct']
def summarize(rep, dst, note):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
    want = WANT + [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    with open(dst, 'w') as f:
        f.write(f'# {note}\nmetric,unit,value\n')
        for w in want:
            if w in d:
                f.write(f'{w},{d[w][0]},"{d[w][1]}"\n')
    def num(k):
        u, v = d[k]
        return float(v) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
    return num('dram__bytes_read.sum'), num('dram__bytes_write.sum'), d['gpu__time_duration.sum'], d['smsp__inst_executed.sum'][1], d['smsp__issue_active.avg.pct_of_peak_sustained_active'][1]
r = summarize(f'gpurun_out/prof_{tag}_c2.ncu-rep', f'profiles/{out}_ncu_full_summary_c2.csv', 'ncu --set full --clock-control none -k regex:v3_kernel -c 1, python bench.py --steps 1 --warmup 3 --profile (c2: 14062 full frames, 24-bit/96 kHz)')
print(r)
json.dump({"kernel": "zf::v3::zf_encode_stereo_v3_kernel<3>", "source": f"profiles/{out}_ncu_full_summary_c2.csv (ncu --set full, one launch, bench c2)", "dram_bytes_read": r[0], "dram_bytes_write": r[1], "dram_bytes_per_launch": r[0] + r[1]}, open(f'profiles/{out}_roofline.json', 'w'))
r = summarize(f'gpurun_out/prof_{tag}_c3.ncu-rep', f'profiles/{out}_ncu_full_summary_c3.csv', 'ncu --set full --clock-control none -k regex:v3_kernel -c 1, python bench.py --workload c3_32bit_192k_600s --steps 1 --warmup 3 --profile (c3: 28125 full frames, 32-bit/192 kHz)')
print(r)
