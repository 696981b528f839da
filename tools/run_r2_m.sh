set -x
python examples/evidence_ladder.py --kmax 4 --epochs 300 --nlive-per-dim 25 --sampler device > gpurun_out/r2m_ladder_device.log 2> gpurun_out/r2m_ladder_device.err; echo rc=$?; cat gpurun_out/r2m_ladder_device.log; tail -3 gpurun_out/r2m_ladder_device.err
python examples/evidence_ladder.py --kmax 3 --epochs 300 --nlive-per-dim 25 --sampler host > gpurun_out/r2m_ladder_host.log 2> gpurun_out/r2m_ladder_host.err; echo rc=$?; cat gpurun_out/r2m_ladder_host.log; tail -3 gpurun_out/r2m_ladder_host.err
