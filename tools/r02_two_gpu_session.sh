#!/bin/bash
# 2-GPU evidence (run through `gpurun --gpus 2`): in-process multi-device tests, the N=2 bench lines, NVLink proof.
mkdir -p gpurun_out/r02_2gpu
O=gpurun_out/r02_2gpu
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -rA -k "in_process_multi_device" > $O/pytest_in_process_multi_device.log 2>&1; tail -4 $O/pytest_in_process_multi_device.log
python tools/r02_nvlink_probe.py 200 > $O/nvlink_counters.txt 2>&1; cat $O/nvlink_counters.txt
ncu --query-metrics --chip gb100 2>/dev/null | grep -i -E "nvl|peer" > $O/ncu_nvlink_metric_names.txt; wc -l $O/ncu_nvlink_metric_names.txt
timeout 600 ncu --devices 1 --metrics "nvltx__bytes.sum,nvlrx__bytes.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_protocol.sum,syslts__t_sectors_aperture_peer.sum,syslts__t_sectors_aperture_peer_op_write.sum,syslts__t_sectors_aperture_peer_op_read.sum,syslts__t_sectors_srcunit_tex_aperture_peer_op_write.sum,syslts__t_requests_srcunit_tex_aperture_peer_op_write.sum,gpu__time_duration.sum,smsp__inst_executed_op_global_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum" \
  --clock-control none -k regex:render_kernel --launch-skip 2 -c 2 --csv --log-file $O/ncu_peer_store.csv python tools/r02_nvlink_probe.py 4 > $O/ncu_peer_store.log 2>&1; tail -3 $O/ncu_peer_store.log; head -c 1500 $O/ncu_peer_store.csv
for mode in fused nccl; do
  RT_BENCH_GATHER=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_2gpu_$mode.json 2> $O/bench_2gpu_$mode.err; tail -c 400 $O/bench_2gpu_$mode.err
done
