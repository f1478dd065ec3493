"""CPU: the profile tooling parses what ncu writes (tools/launch_summary.py feeds profiles/*_summary.csv)."""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent

NCU_CSV = '''==PROF== Connected to process 1
"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"
"0","1","python","h","void ltxv::<unnamed>::gemm_bf16_tn_kernel<192>(CUtensorMap_st, CUtensorMap_st, ltxv::GemmParams)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_read.sum","Mbyte","100.5"
"0","1","python","h","void ltxv::<unnamed>::gemm_bf16_tn_kernel<192>(CUtensorMap_st, CUtensorMap_st, ltxv::GemmParams)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","dram__bytes_write.sum","Mbyte","10"
"0","1","python","h","void ltxv::<unnamed>::gemm_bf16_tn_kernel<192>(CUtensorMap_st, CUtensorMap_st, ltxv::GemmParams)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","90"
"1","1","python","h","void ltxv::<unnamed>::gemm_bf16_tn_kernel<192>(CUtensorMap_st, CUtensorMap_st, ltxv::GemmParams)","1","7","(256, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","us","110"
"2","1","python","h","ltxv::<unnamed>::flash_attn3_kernel(CUtensorMap_st, CUtensorMap_st, CUtensorMap_st, ltxv::AttnParams, ltxv::<unnamed>::SplitPlan)","1","7","(384, 1, 1)","(640, 1, 1)","0","10.0","Command line profiler metrics","gpu__time_duration.sum","ms","0.3"
'''


def test_launch_summary_groups_kernels_and_units(tmp_path):
    f = tmp_path / "launches.csv"
    f.write_text(NCU_CSV)
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), str(f), "unit test"],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    assert out[0] == "# unit test"
    assert out[1] == "# total 500 us over 3 launches"
    rows = [r.split(",") for r in out[3:]]
    assert rows[0][:4] == ["1", "300.0", "300.0", "60.0"] and rows[0][-1] == "flash_attn3_kernel"
    assert rows[1][:6] == ["2", "200.0", "100.0", "40.0", "50.2", "5.0"] and rows[1][-1] == "gemm_bf16_tn_kernel<192>"
