#!/bin/bash
# BASELINE.json configs[4] on ONE 8-GPU box:
#   A. the full 3.8 M-segment corpus (118 750 ragged clips) on 8 GPUs;
#   B. a 1/16 subset (7 422 clips, ~237 k segments = the corpus's first clips) on 1, 2 and 4 GPUs -- concurrently, on
#      disjoint GPUs -- and then on 8;
#   C. bit-identity of the per-clip results across all runs (subset runs vs each other and vs the full run's prefix).
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --nproc-per-node 8 --master-port 29511 tools/run_corpus.py --clips 118750 --out gpurun_out/corpus_full_n8.json 2> gpurun_out/corpus_full_n8.err | tail -1
SUB=7422
CUDA_VISIBLE_DEVICES=0 python tools/run_corpus.py --clips $SUB --out gpurun_out/corpus_sub_n1.json 2> gpurun_out/corpus_sub_n1.err | tail -1 &
CUDA_VISIBLE_DEVICES=1,2 $TR --nproc-per-node 2 --master-port 29512 tools/run_corpus.py --clips $SUB --out gpurun_out/corpus_sub_n2.json 2> gpurun_out/corpus_sub_n2.err | tail -1 &
CUDA_VISIBLE_DEVICES=3,4,5,6 $TR --nproc-per-node 4 --master-port 29513 tools/run_corpus.py --clips $SUB --out gpurun_out/corpus_sub_n4.json 2> gpurun_out/corpus_sub_n4.err | tail -1 &
wait
$TR --nproc-per-node 8 --master-port 29514 tools/run_corpus.py --clips $SUB --out gpurun_out/corpus_sub_n8.json 2> gpurun_out/corpus_sub_n8.err | tail -1
python tools/run_corpus.py --compare gpurun_out/corpus_sub_n1.json gpurun_out/corpus_sub_n2.json gpurun_out/corpus_sub_n4.json gpurun_out/corpus_sub_n8.json gpurun_out/corpus_full_n8.json | tee gpurun_out/corpus_compare.json
