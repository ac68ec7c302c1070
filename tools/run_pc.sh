free -g | head -2
CONFIGS=4big M_BIG=10000000 timeout 900 python tools/configs.py 2>gpurun_out/big.err | tee gpurun_out/r1_config_10M.jsonl; tail -3 gpurun_out/big.err
