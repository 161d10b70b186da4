#!/bin/bash
# One `ncu --set full` capture of one kernel launch, reduced on the box to text:
#   gpurun_out/<name>_metrics.txt   selected raw metrics (time, DRAM bytes, pipes, stalls, L1/LSU, occupancy)
#   gpurun_out/<name>_hotlines.txt  top source lines by warp-stall samples (tools/ncu_lines.py)
# usage: tools/ncu_cap.sh NAME KERNEL_REGEX SKIP command...
name=$1; rx=$2; skip=$3; shift 3
mkdir -p gpurun_out /tmp/ncu
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -f -o /tmp/ncu/$name "$@" > /tmp/ncu/$name.log 2>&1
python - "$name" "$*" <<'PY'
import csv, re, subprocess, sys
name, cmd = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", f"/tmp/ncu/{name}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
keep = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__throughput.avg.pct|inst_executed.avg.per_cycle_elapsed|"
                  r"smsp__inst_executed.sum$|thread_inst_executed_per_inst|warps_active.avg.pct|launch__(registers|grid_size|block_size|occupancy_limit|shared_mem_per_block)|"
                  r"l1tex__data_pipe_lsu_wavefronts|l1tex__throughput.avg.pct|lts__throughput.avg.pct|hit_rate.pct|bank_conflicts|issue_active.avg.pct|"
                  r"smsp__average_warps?_(latency_)?issue_stalled.*_per_issue_active|smsp__pcsamp_warps_issue_stalled|sm__inst_executed_pipe_.*(sum|pct)|"
                  r"smsp__inst_executed_op_.*sum$|l1tex__t_(requests|sectors)_pipe_lsu_mem_(global|local)_op_(ld|st).sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.*sum$|"
                  r"sm__pipe_(alu|fma|fmaheavy|xu|lsu).*pct|smsp__inst_issued.*per_cycle|sm__cycles_elapsed.avg$|l1tex__lsu_writeback|l1tex__f_wavefronts|lsu_mem_global_op")
with open(f"gpurun_out/{name}_metrics.txt", "w") as f:
    f.write(f"# {name}: ncu --set full --clock-control none, one launch of: {cmd}\n")
    if len(rows) >= 3:
        hdr, units, vals = rows[0], rows[1], rows[2]
        for h, u, v in zip(hdr, units, vals):
            if keep.search(h):
                f.write(f"{h:100s} {v} {u}\n")
    else:
        f.write("capture failed\n" + open(f"/tmp/ncu/{name}.log").read()[-2000:])
PY
{ echo "# $name: top source lines by warp-stall samples (command: $*)"; python tools/ncu_lines.py /tmp/ncu/$name.ncu-rep 45; } > gpurun_out/${name}_hotlines.txt 2>&1
rm -f /tmp/ncu/$name.ncu-rep
