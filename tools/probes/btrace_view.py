"""Print a few rows of the backward kernel's clock trace (tools/probes/run_btrace.sh) on a common time base."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tc_btrace.txt")]
names = ["free", "Afull", "Asdp", "pre", "hat", "s_pre", "s_sdp", "s_sfree", "s_comp", "s_staged", "B_staged", "B_ready", "B_issued",
         "e_acc", "e_accfree", "e_empty"]
base = rows[20][0]
for k in range(20, 25):
    print(k, " ".join(f"{n}={v - base}" for n, v in zip(names, rows[k])))
print("period", (rows[40][0] - rows[20][0]) / 20)
