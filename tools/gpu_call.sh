#!/bin/bash
# tools/gpu_call.sh — the command list of ONE gpurun call, as named stages (what each call measured is summarised in
# profiles/README.md).  Everything it writes goes to gpurun_out/ (merged back by gpurun).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_call.sh tests sweeps ncu'
# stages: final ncufinal golden tests shims sweeps awq refbench refstamps ncu ncufull bench translate pdlrace trprofile smemsweep widetiles awqtrace probe ncutr tp2
set -u
cd "$(dirname "$0")/.."
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/box.txt; nproc >> $OUT/box.txt
M8=/tmp/ct2b200_bench/llama_8b_int8_float16
MA=/tmp/ct2b200_bench/llama_8b_awq_gemm

stage_golden() {   # the reference's CUDA kernels on seeded inputs -> fixtures (needs oracle/_ref_cuda)
  timeout 600 python tools/ref_cuda_worker.py awq-golden $OUT/awq_ref_cuda.npz > $OUT/ref_golden.log 2>&1
  timeout 600 python tools/ref_cuda_worker.py dense-s8 $OUT/dense_s8_ref_cuda.npz >> $OUT/ref_golden.log 2>&1
  cp $OUT/awq_ref_cuda.npz $OUT/dense_s8_ref_cuda.npz tests/golden/ 2>/dev/null
}

stage_tests() {    # the opt-in kernel first, bounded; then the whole GPU suite with the defaults of the tree
  CT2B200_AWQ_GEMV=1 timeout 600 python -m pytest tests/test_gpu_awq.py tests/test_gpu_ref_cuda.py -q -k "awq" > $OUT/pytest_awq_gemv.log 2>&1
  echo "awq gemv tests exit $?" >> $OUT/pytest_awq_gemv.log
  timeout 1500 python -m pytest tests -m gpu -q --tb=short > $OUT/pytest_gpu.log 2>&1
  echo "gpu suite exit $?" >> $OUT/pytest_gpu.log
}

stage_shims() {    # the reference's own gtests with libct2b200 interposed under its ops (oracle/Makefile.shims)
  CT2B200_SHIM_REPORT=$OUT/shim_report.txt timeout 600 oracle/_ref_cuda/ct2_tests_b200 tests/golden \
    --gtest_filter='*CUDA*' > $OUT/ref_gtests_on_b200.txt 2>&1
  echo "shim gtests exit $?" >> $OUT/ref_gtests_on_b200.txt
}
stage_translate() {   # encoder-decoder path: tests first, then the OPUS-MT-shaped bench record
  timeout 900 python -m pytest tests/test_gpu_translator.py tests/test_gpu_whisper.py -q --tb=short > $OUT/pytest_translator.log 2>&1
  echo "translator tests exit $?" >> $OUT/pytest_translator.log
}

run() { echo "== B=$B $*" >> $OUT/sweep.log; env "$@" timeout 300 python tools/decode_once.py $B 64 int8_float16 8b int8_float16 >> $OUT/sweep.log 2>&1; }
stage_sweeps() {   # decode step of the INT8 8B model (64 steps after the 1024-token prompt)
  for B in 1 32; do
    run CT2B200_PDL=1
  done
}
awq_run() { echo "== AWQ batch=$1 CT2B200_AWQ_DECODE=$2 CT2B200_AWQ_GEMV=$3" >> $OUT/sweep.log
  CT2B200_AWQ_DECODE=$2 CT2B200_AWQ_GEMV=$3 timeout 600 python tools/decode_once.py $1 64 float16 8b awq_gemm >> $OUT/sweep.log 2>&1; }
stage_awq() {
  awq_run 1 1 1; awq_run 32 1 0; awq_run 8 1 0
  echo "== AWQ batch=32, full-height tiles pinned" >> $OUT/sweep.log
  CT2B200_GEMM_ROWS=128 timeout 600 python tools/decode_once.py 32 64 float16 8b awq_gemm >> $OUT/sweep.log 2>&1
}

stage_refbench() { # the reference's CUDA build on the same workload (bounded: 16 / 80 generated tokens)
  python -c "import bench; bench.model_dir('8b','int8_float16'); bench.model_dir('8b','awq_gemm')" 2>> $OUT/ref_cuda_bench.log
  for b in 1 32; do
    timeout 600 python tools/ref_cuda_worker.py bench $M8 int8_float16 $b 1024 16 80 >> $OUT/ref_cuda_bench.log 2>&1
    timeout 600 python tools/ref_cuda_worker.py bench $M8 int8_float16 $b 1024 16 80 --flash >> $OUT/ref_cuda_bench.log 2>&1
    timeout 600 python tools/ref_cuda_worker.py bench $MA float16 $b 1024 16 80 >> $OUT/ref_cuda_bench.log 2>&1
    timeout 600 python tools/ref_cuda_worker.py bench $MA float16 $b 1024 16 80 --flash >> $OUT/ref_cuda_bench.log 2>&1
  done
}

stage_ncu() {      # launch lists of the INT8 step (2 steps) at bsz 32 and 1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_b32.csv python tools/decode_once.py 32 2 int8_float16 8b int8_float16 > $OUT/ncu_list.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_b1.csv python tools/decode_once.py 1 2 int8_float16 8b int8_float16 >> $OUT/ncu_list.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_awq_b1.csv env CT2B200_AWQ_DECODE=1 python tools/decode_once.py 1 2 float16 8b awq_gemm >> $OUT/ncu_list.log 2>&1
}
stage_ncufull() {  # full captures of the AWQ decode kernel and of the INT8 decode GEMM (qkv / out / gate+up / down of two layers)
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:awq_decode_kernel -c 6 \
    -o $OUT/r02_awq_decode_v3 python tools/decode_once.py 32 2 float16 8b awq_gemm > $OUT/ncu_awq.log 2>&1
  timeout 300 python tools/ncu_extract.py $OUT/r02_awq_decode_v3.ncu-rep > $OUT/r02_ncu_awq_v3.md 2>> $OUT/ncu_awq.log
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_decode_kernel -c 6 \
    -o $OUT/r02_gemm_decode_v3 python tools/decode_once.py 32 2 int8_float16 8b int8_float16 > $OUT/ncu_gemm.log 2>&1
  timeout 300 python tools/ncu_extract.py $OUT/r02_gemm_decode_v3.ncu-rep > $OUT/r02_ncu_gemm_v3.md 2>> $OUT/ncu_gemm.log
}
stage_bench() {
  ( time timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time
  echo "bench exit $?" >> $OUT/bench.err
}

stage_pdlrace() {  # eager decode loop under programmatic dependent launch: product build vs a build without ld.global.nc
  T=tests/test_gpu_engine.py::test_generate_scores_match_reference
  for i in 1 2 3; do
    echo "=== product build, run $i" >> $OUT/pdlrace.log
    timeout 300 python -m pytest $T -q --tb=line 2>&1 | tail -3 >> $OUT/pdlrace.log
    echo "=== norestrict build (no read-only loads), run $i" >> $OUT/pdlrace.log
    CT2B200_LIB=$PWD/ctranslate2_b200/libct2b200_norestrict.so timeout 300 python -m pytest $T -q --tb=line 2>&1 | tail -3 >> $OUT/pdlrace.log
  done
  echo "=== compute-sanitizer memcheck, eager" >> $OUT/pdlrace.log
  timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest "$T[False]" -q --tb=line 2>&1 | tail -25 >> $OUT/pdlrace.log
}
stage_refstamps() { # the reference's CUDA build: decode ms/step from the per-step callback stamps
  python -c "import bench; bench.model_dir('8b','int8_float16'); bench.model_dir('8b','awq_gemm')" 2>> $OUT/ref_cuda_stamps.log
  for b in 1 32; do
    timeout 600 python tools/ref_cuda_worker.py bench $M8 int8_float16 $b 1024 8 48 >> $OUT/ref_cuda_stamps.log 2>&1
    timeout 600 python tools/ref_cuda_worker.py bench $MA float16 $b 1024 8 48 >> $OUT/ref_cuda_stamps.log 2>&1
  done
}
stage_smemsweep() { # does a smaller operand ring (successor GEMM co-resident earlier) help the INT8 step?
  for B in 1 32; do
    run CT2B200_PDL=1
    run CT2B200_GEMM_SMEM_KB=112
    run CT2B200_GEMM_SMEM_KB=96
    run CT2B200_GEMM_SMEM_KB=64
  done
}
stage_trprofile() {   # launch list of the OPUS-MT-shaped decoding step + timing
  timeout 300 python tools/translate_once.py 64 4 64 > $OUT/translate_once.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_translate.csv python tools/translate_once.py 64 4 2 >> $OUT/translate_once.log 2>&1
}

stage_widetiles() { # the decode GEMM with 128 / 256 activation rows (opt-in): parity, then the translation step with and without it
  CT2B200_GEMM_DECODE_MAXM=256 timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_translator.py -q -k "dense or translat or golden or reference" --tb=short > $OUT/pytest_widetiles.log 2>&1
  echo "wide tiles tests exit $?" >> $OUT/pytest_widetiles.log
  timeout 300 python tools/translate_once.py 64 4 64 > $OUT/translate_widetiles.log 2>&1
  CT2B200_GEMM_DECODE_MAXM=256 timeout 300 python tools/translate_once.py 64 4 64 >> $OUT/translate_widetiles.log 2>&1
  CT2B200_GEMM_DECODE_MAXM=128 timeout 300 python tools/translate_once.py 64 4 64 >> $OUT/translate_widetiles.log 2>&1
  CT2B200_GEMM_DECODE_MAXM=256 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_translate_wide.csv python tools/translate_once.py 64 4 2 >> $OUT/translate_widetiles.log 2>&1
}
stage_awqtrace() { # pipeline stamps of the AWQ decode kernel (trace build)
  CT2B200_LIB=$PWD/ctranslate2_b200/libct2b200_awqtrace.so timeout 300 python tools/awq_trace.py 32 > $OUT/awq_trace.log 2>&1
}
stage_ncutr() {     # full capture of the two slowest kernels of the translation step
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"beam_rows" -c 2 \
    -o $OUT/r02_translate_kernels python tools/translate_once.py 64 4 2 > $OUT/ncu_tr.log 2>&1
  timeout 300 python tools/ncu_extract.py $OUT/r02_translate_kernels.ncu-rep > $OUT/r02_ncu_translate.md 2>> $OUT/ncu_tr.log
}
stage_probe() {    # ingest rate of one CTA per SM as a function of the TMA request shape (tools/probes/tma_probe.cu)
  [ -x tools/probes/tma_probe ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/tma_probe tools/probes/tma_probe.cu -lcuda
  timeout 120 tools/probes/tma_probe > $OUT/tma_probe.log 2>&1
}
stage_tpcheck() {   # needs gpurun --gpus 2: the parity test, then the worker three more times (the upload race was intermittent)
  timeout 600 python -m pytest tests/test_gpu_tp.py -q --tb=short > $OUT/pytest_tp.log 2>&1
  echo "tp tests exit $?" >> $OUT/pytest_tp.log
  for i in 1 2 3; do
    echo "=== repeat $i" >> $OUT/tpdebug.log
    TP_WORKER_REPORT_ONLY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29517 tests/tp_worker.py 2>&1 | grep "TP_CASE\|Error\|error" >> $OUT/tpdebug.log
  done
}
stage_tpdebug() {   # needs gpurun --gpus 2: the tensor-parallel parity worker under the switches that changed this round
  for e in A=0 CT2B200_GEMM_PREFILL_BN=256 CT2B200_GEMM_PREFILL=0 CT2B200_PDL=0 CT2B200_GEMM_DECODE=0; do
    echo "=== $e" >> $OUT/tpdebug.log
    env $e TP_WORKER_REPORT_ONLY=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29517 tests/tp_worker.py 2>&1 | grep "TP_CASE\|Error\|error" >> $OUT/tpdebug.log
  done
}
stage_tp2() {      # needs gpurun --gpus 2: tensor-parallel parity (tests/tp_worker.py) and the bench line with its `tp` record
  timeout 900 python -m pytest tests/test_gpu_tp.py -q --tb=short > $OUT/pytest_tp.log 2>&1
  echo "tp tests exit $?" >> $OUT/pytest_tp.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 256 --warmup 3 > $OUT/bench_gpus2.json 2> $OUT/bench_gpus2.err
  echo "bench --gpus 2 exit $?" >> $OUT/bench_gpus2.err
}

stage_final() {    # last call of the round (bounded to the GPU minutes left): the whole suite like the driver runs it, smoke(),
                   # the launch list of the final kernels, then the default bench line with a short side budget
  ( time timeout 480 python -m pytest tests -m gpu -q --tb=short --durations=8 > $OUT/pytest_gpu.log 2>&1 ) 2> $OUT/pytest_gpu.time
  echo "gpu suite exit $?" >> $OUT/pytest_gpu.log
  timeout 120 python __graft_entry__.py --smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log
  timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_b32_final.csv python tools/decode_once.py 32 2 int8_float16 8b int8_float16 > $OUT/ncu_list.log 2>&1
  ( time timeout 420 python bench.py --side-budget 100 > $OUT/bench.json 2> $OUT/bench.err ) 2> $OUT/bench.time
  echo "bench exit $?" >> $OUT/bench.err
}

stage_ncufinal() { # full captures of the final kernels (INT8 decode GEMMs of two layers, decode attention, AWQ decode kernel) + launch lists
  timeout 170 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_decode_kernel -c 8 \
    -o $OUT/r02_final_gemm_decode python tools/decode_once.py 32 2 int8_float16 8b int8_float16 > $OUT/ncu_gemm.log 2>&1
  timeout 60 python tools/ncu_extract.py $OUT/r02_final_gemm_decode.ncu-rep > $OUT/r02_ncu_final_gemm_decode.md 2>> $OUT/ncu_gemm.log
  timeout 120 ncu --set full --clock-control none --profile-from-start off -k regex:attention_decode -c 2 \
    -o $OUT/r02_final_attention python tools/decode_once.py 32 2 int8_float16 8b int8_float16 > $OUT/ncu_attn.log 2>&1
  timeout 60 python tools/ncu_extract.py $OUT/r02_final_attention.ncu-rep > $OUT/r02_ncu_final_attention.md 2>> $OUT/ncu_attn.log
  timeout 170 ncu --set full --clock-control none --profile-from-start off -k regex:awq_decode_kernel -c 6 \
    -o $OUT/r02_final_awq_decode python tools/decode_once.py 32 2 float16 8b awq_gemm > $OUT/ncu_awq.log 2>&1
  timeout 60 python tools/ncu_extract.py $OUT/r02_final_awq_decode.ncu-rep > $OUT/r02_ncu_final_awq_decode.md 2>> $OUT/ncu_awq.log
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_awq_b32_final.csv python tools/decode_once.py 32 2 float16 8b awq_gemm > $OUT/ncu_list.log 2>&1
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/r02_launches_b1_final.csv python tools/decode_once.py 1 2 int8_float16 8b int8_float16 >> $OUT/ncu_list.log 2>&1
  ls -la $OUT/*.ncu-rep >> $OUT/ncu_list.log 2>&1
}

for s in "$@"; do
  echo "=== stage $s $(date +%T)" >> $OUT/stages.log
  stage_$s
done
echo "=== done $(date +%T)" >> $OUT/stages.log
tail -3 $OUT/sweep.log 2>/dev/null; tail -3 $OUT/pytest_gpu.log 2>/dev/null
