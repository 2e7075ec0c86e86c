"""Extract the judged subset of an `ncu --page raw --csv` dump into a small metric,value,unit table.
usage: python tools/ncu_key_metrics.py gpurun_out/TAG/prof_fused_raw.csv > profiles/rNN/ncu_..._key_metrics.csv"""
import csv
import re
import sys

KEEP = re.compile(
    r"^(Kernel Name|Block Size|Grid Size|dram__bytes_(read|write)\.sum.*|gpu__dram_throughput.*|gpu__time_duration\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum.*|"
    r"launch__(occupancy_limit_shared_mem|registers_per_thread.*|shared_mem_per_block_dynamic)|"
    r"sm__cycles_elapsed\.max.*|sm__inst_executed_pipe_(tc|uniform|tensor_subpipe_\w+)\.(avg|max|min|sum)\.pct.*|"
    r"sm__issue_active\.avg\.pct.*|sm__mem_tensor_cycles_active\.avg.*|sm__pipe_(fma|tensor|xu|alu)_cycles_active\.avg.*|"
    r"sm__throughput\.avg.*|sm__warps_active\.avg.*|smsp__average_warps_issue_stalled_.*|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct.*)$")


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = csv.writer(sys.stdout, quoting=csv.QUOTE_ALL)
    sys.stdout.write("metric,value,unit\n")
    for h, u, v in sorted(zip(hdr, units, vals), key=lambda t: (not t[0][0].isupper(), t[0])):
        if KEEP.match(h):
            out.writerow([h, v, u])


if __name__ == "__main__":
    main(sys.argv[1])
