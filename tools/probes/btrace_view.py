"""Print a few rows of the backward kernel's clock trace (HV_TC_BTRACE_DUMP of a -DHV_TC_TRACE build) on a common time base."""
import sys
rows = [list(map(int, l.split())) for l in open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tc_btrace.txt")]
names = ["free", "Afull", "Asdp", "-", "-", "s_full", "s_sdp", "s_sfree", "s_D", "s_staged", "B_staged", "-", "B_issued",
         "e_acc", "st_written", "st_released"]
base = rows[20][0]
for k in range(20, 26):
    print(k, " ".join(f"{n}={v - base}" for n, v in zip(names, rows[k]) if n != "-"))
print("period", (rows[40][0] - rows[20][0]) / 20)
