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


def test_gemm_traffic_json_is_reproducible_from_the_committed_launch_list():
    """bench.py's roofline.traffic comes from profiles/r02e_gemm_traffic.json; that file must be exactly what
    `launch_summary.py --traffic` derives from the committed ncu launch list (no hand-edited numbers)."""
    import json
    csv_path = ROOT / "profiles" / "r02e_launches_step_plus_decode.csv"
    committed = json.loads((ROOT / "profiles" / "r02e_gemm_traffic.json").read_text())
    out = subprocess.run([sys.executable, str(ROOT / "tools" / "launch_summary.py"), "--traffic",
                          str(csv_path.relative_to(ROOT))], capture_output=True, text=True, check=True, cwd=str(ROOT)).stdout
    derived = json.loads(out)
    for k in ("dit_step_launches", "gemm_launches"):
        assert derived[k] == committed[k]
    for k in ("bytes_per_launch_mean", "gemm_share_of_step", "dit_step_us_under_ncu"):
        assert abs(derived[k] - committed[k]) <= 1e-9 * abs(committed[k])
    assert set(derived["per_kernel"]) == set(committed["per_kernel"])
    # the denoise step of the capture: the batched-CFG forward's GEMMs run at M = 9984 behind 28 attention launches
    assert committed["gemm_launches"] >= 6 * 28 and 0.5 < committed["gemm_share_of_step"] < 0.7
    import bench
    assert abs(bench.ncu_traffic_bytes() - committed["bytes_per_launch_mean"]) < 1.0
