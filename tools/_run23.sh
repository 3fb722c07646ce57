mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/launches_bench.log 2>&1
cap() { # name regex config skip
  timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $4 -c 1 -f -o gpurun_out/ncu_r2_final_$1 python bench.py --config $3 --steps 3 --warmup 3 --no-extras > /dev/null 2>&1
  ls -la gpurun_out/ncu_r2_final_$1.ncu-rep
}
cap slim gbdt_score_slim C2 4
cap rowgather "row_gather|code_gather" C2 4
cap cosine cosine_f32 C4 4
cap leaves gbdt_leaves C5 4
cap sum gbdt_sum C5 4
cap assemble "assemble_kernel" C3 4
cap prepass "prepass_kernel" C3 4
timeout 300 ncu --set full --clock-control none --import-source on -k regex:encoder_gemm -s 2 -c 1 -f -o gpurun_out/ncu_r2_final_gemm python tools/gemm_prof.py 262144 1152 384 > /dev/null 2>&1; ls -la gpurun_out/ncu_r2_final_gemm.ncu-rep
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_short -s 2 -c 1 -f -o gpurun_out/ncu_r2_final_att python tools/enc_prof.py 16384 16 > /dev/null 2>&1; ls -la gpurun_out/ncu_r2_final_att.ncu-rep
