"""Summarise .ncu-rep captures (ncu --set full) into one JSON: python profiles/ncu_summary.py out.json name=file.ncu-rep ..."""
import csv, io, json, subprocess, sys
WANT = {
    "gpu__time_duration.sum": "duration",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__block_size": "block_size", "launch__grid_size": "grid_size",
    "launch__shared_mem_per_block_dynamic": "dynamic_smem_per_block",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1tex__data_pipe_lsu_wavefronts_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "shared_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum": "shared_atomic_wavefronts",
    "smsp__inst_executed_op_shared_atom.sum": "shared_atomic_instructions",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "shared_bank_conflicts",
    "lts__t_sectors_op_red.sum": "l2_red_sectors", "lts__t_sectors_op_atom.sum": "l2_atom_sectors",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
}
out = {}
for arg in sys.argv[2:]:
    name, path = arg.split("=", 1)
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")][:160]}
        for i, h in enumerate(hdr):
            if h in WANT and vals[i] != "":
                try:
                    d[WANT[h]] = float(vals[i].replace(",", ""))
                except ValueError:
                    d[WANT[h]] = vals[i]
                d[WANT[h] + "_unit"] = units[i]
        launches.append(d)
    out[name] = launches
json.dump(out, open(sys.argv[1], "w"), indent=1)
print("wrote", sys.argv[1], {k: len(v) for k, v in out.items()})
