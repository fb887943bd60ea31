"""Turns the artefacts of tools/gpu_full_run.sh (gpurun_out/<tag>_*) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <tag> <out-prefix>

Writes profiles/<out>_bench*.json, <out>_ncu_launches.csv, <out>_ncu_full_summary_{c1,c2,c3}.csv and <out>_roofline.json
(`captures`: per workload the DRAM bytes, executed warp instructions and issue-slot utilisation of one launch of the dominant
kernel -- what bench.py reports as roofline.traffic / roofline.issue for that workload).
"""
import csv, json, os, shutil, subprocess, sys
tag, out = sys.argv[1], sys.argv[2]
for f in ['bench', 'bench_c1', 'bench_c3', 'bench_c4', 'bench_ref', 'bench_c5']:
    src = f'gpurun_out/{tag}_{f}.json'
    if not os.path.exists(src):
        continue
    try:
        d = json.loads(open(src).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, 'unreadable', e)
        continue
    print(f, d['value'], d['ms_per_step'], (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('kernel_ms'), d['e2e']['value'], d.get('gpu_launches'))
    shutil.copy(src, f'profiles/{out}_{f}.json')
if os.path.exists(f'gpurun_out/{tag}_launches.csv'):
    shutil.copy(f'gpurun_out/{tag}_launches.csv', f'profiles/{out}_ncu_launches.csv')
WANT = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second', 'dram__bytes_write.sum.per_second', 'smsp__sass_inst_executed_op_local_ld.sum', 'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__icc_request_hit_rate.pct', 'sm__cycles_elapsed.max']
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
        return float(v.replace(',', '')) * {'Mbyte': 1e6, 'Gbyte': 1e9, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
    return {"kernel": d['Kernel Name'][1], "dram_bytes_read": num('dram__bytes_read.sum'), "dram_bytes_write": num('dram__bytes_write.sum'),
            "dram_bytes_per_launch": num('dram__bytes_read.sum') + num('dram__bytes_write.sum'),
            "inst_executed": num('smsp__inst_executed.sum'), "issue_active_pct": num('smsp__issue_active.avg.pct_of_peak_sustained_active'),
            "duration": ' '.join(d['gpu__time_duration.sum'][::-1]),
            "shared_bank_conflicts": num('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') if 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum' in d else None}
WL = {'c1': ('c1_16bit_44k1_60s', '--workload c1_16bit_44k1_60s '), 'c2': ('c2_24bit_96k_600s', ''), 'c3': ('c3_32bit_192k_600s', '--workload c3_32bit_192k_600s ')}
caps = []
for key, (wl, arg) in WL.items():
    rep = f'gpurun_out/prof_{tag}_{key}.ncu-rep'
    if not os.path.exists(rep):
        continue
    dst = f'profiles/{out}_ncu_full_summary_{key}.csv'
    r = summarize(rep, dst, f'ncu --set full --clock-control none --import-source on -k regex:v3_kernel -c 1, python bench.py {arg}--steps 1 --warmup 3 --profile')
    r.update(workload=wl, source=dst + ' (one launch)')
    print(key, r)
    caps.append(r)
if caps:
    json.dump({"captures": caps}, open(f'profiles/{out}_roofline.json', 'w'), indent=1)
# decoder: launch list, one full capture of the frames kernel, the probe's timings
for src, dst in [(f'gpurun_out/{tag}_launches_decode.csv', f'profiles/{out}_ncu_launches_decode.csv'),
                 (f'gpurun_out/{tag}_decode_probe.txt', f'profiles/{out}_decode_probe.txt'),
                 (f'gpurun_out/{tag}_pytest.log', f'profiles/{out}_pytest_gpu.log')]:
    if os.path.exists(src):
        shutil.copy(src, dst)
rep = f'gpurun_out/prof_{tag}_decode.ncu-rep'
if os.path.exists(rep):
    r = summarize(rep, f'profiles/{out}_ncu_full_summary_decode.csv',
                  'ncu --set full --clock-control none --import-source on -k regex:zf_dec_frames -c 1, python tools/decode_probe.py c2 1')
    print('decode', r)
