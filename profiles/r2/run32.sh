mkdir -p gpurun_out
timeout 100 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > gpurun_out/smoke_final3.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final3.log
