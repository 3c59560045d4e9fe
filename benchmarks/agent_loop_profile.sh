# Kernel list of one replay of the captured Test_Agent loop (B = 1 and B = 32) and a full capture of the agent-side
# kernels of this library inside it:   gpurun --timeout 600 -- 'bash benchmarks/agent_loop_profile.sh'
# then   python profiles/make_agent_loop.py
set -x
mkdir -p gpurun_out
for B in 1 32; do
  timeout 120 python benchmarks/debug/agent_loop_launches.py $B > gpurun_out/r2h_loop_b$B.log 2>&1 && \
  timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
      --log-file gpurun_out/r2h_loop_b$B.csv python benchmarks/debug/agent_loop_launches.py $B > gpurun_out/r2h_ncu_b$B.log 2>&1
  echo "loop B=$B exit $?"
done
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"k_grouped_linear|k_conv_epilogue" -c 14 -o gpurun_out/r2h_agent_kernels -f \
    python benchmarks/debug/agent_loop_launches.py 32 > gpurun_out/r2h_ncu_full.log 2>&1
echo "full capture exit $?"
