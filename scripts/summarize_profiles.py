"""Turn an ncu report of the rows kernel into the tracked summaries under profiles/:

    python scripts/summarize_profiles.py gpurun_out/r02f_rows_kernel.ncu-rep r02f_rows_kernel "<how it was captured>"

writes profiles/<name>_raw.csv (`--page raw --csv`), profiles/<name>_source_top.txt (per CUDA source line, through
scripts/ncu_source_lines.py) and profiles/rows_kernel_profile.json -- the numbers bench.py prints as
`roofline.traffic` / `roofline_compute`, stamped with the sha256 of the kernel sources they belong to (bench.py
marks them stale when the sources changed since)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12,
        'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}


def raw_row(rep, out_csv):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    with open(out_csv, 'w') as f:
        f.write(txt)
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {}
    for h, u, v in zip(hdr, units, vals):
        try:
            d[h] = float(v.replace(',', '')) * UNIT.get(u.split('/')[0], 1.0)
        except ValueError:
            d[h] = v
    return d


def main():
    rep, name = sys.argv[1], sys.argv[2]
    how = sys.argv[3] if len(sys.argv) > 3 else ''
    prof = os.path.join(ROOT, 'profiles')
    d = raw_row(rep, os.path.join(prof, name + '_raw.csv'))
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    tmp = os.path.join(prof, name + '_source.csv.tmp')
    with open(tmp, 'w') as f:
        f.write(src)
    top = subprocess.run([sys.executable, os.path.join(ROOT, 'scripts', 'ncu_source_lines.py'), tmp, '45'],
                         capture_output=True, text=True).stdout
    os.remove(tmp)
    with open(os.path.join(prof, name + '_source_top.txt'), 'w') as f:
        f.write(top)
    import bench
    st = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
    out = {
        'src_sha16': bench.src_sha16(),
        'kernel': d.get('Kernel Name', 'rows_kernel'),
        'grid': int(d['launch__grid_size']), 'block': int(d['launch__block_size']),
        'registers': int(d['launch__registers_per_thread']),
        'dyn_smem_kb': d['launch__shared_mem_per_block_dynamic'] / 1e3,
        'duration_ms_under_ncu': d['gpu__time_duration.sum'],
        'dram_bytes_per_launch': d['dram__bytes_read.sum'] + d['dram__bytes_write.sum'],
        'dram_read_bytes': d['dram__bytes_read.sum'], 'dram_write_bytes': d['dram__bytes_write.sum'],
        'fp64_pipe_pct_active': d['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active'],
        'issue_active_pct': d['sm__inst_issued.avg.pct_of_peak_sustained_active'],
        'inst_per_cycle_active': d['sm__inst_executed.avg.per_cycle_active'],
        'stall_barrier': d[st % 'barrier'], 'stall_wait': d[st % 'wait'],
        'stall_long_scoreboard': d[st % 'long_scoreboard'], 'stall_short_scoreboard': d[st % 'short_scoreboard'],
        'l2_hit_pct': d['lts__t_sector_hit_rate.pct'],
        'warp_instructions': int(d['smsp__inst_executed.sum']),
        'source': 'profiles/%s_raw.csv (%s)' % (name, how),
    }
    with open(os.path.join(prof, 'rows_kernel_profile.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))
    print(top[:3000])


if __name__ == '__main__':
    main()
