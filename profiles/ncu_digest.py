"""Digest of one `ncu --set full` report: headline metrics + the instructions with the most stall samples.
usage: python profiles/ncu_digest.py report.ncu-rep [n_top=24]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__inst_executed.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']


def page(rep, name):
    out = subprocess.run(['ncu', '-i', rep, '--page', name, '--csv'], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, n_top=24):
    rows = page(rep, 'raw')
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(f"== {rep}: {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:90s} {units[i]:12s} {vals[i]}")
    rows = page(rep, 'source')
    hdr, data = rows[1], rows[2:]
    i_s, i_src, i_ex = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot = sum(int(r[i_s]) for r in data)
    print(f"-- warp stall samples: {tot}; top instructions (SASS index, instruction, samples, executions, top stall reasons)")
    for i in sorted(sorted(range(len(data)), key=lambda i: -int(data[i][i_s]))[:n_top]):
        r = data[i]
        st = sorted(((hdr[j][6:], int(r[j])) for j in stall if int(r[j]) > 0), key=lambda x: -x[1])[:3]
        print(f"{i:5d} {r[i_src].strip()[:64]:64s} {r[i_s]:>6s} {r[i_ex]:>9s} {st}")


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
