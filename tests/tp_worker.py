"""One rank of the tensor-parallel parity test (launched by torchrun, one process per GPU; see test_gpu_tp.py).
Checks, on every rank: TP logits == single-GPU logits (same engine, tp_size 1) within fp tolerance, greedy tokens
identical across ranks and equal to the single-GPU ones, and the oracle bound of the 2-layer model."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctranslate2_b200 as ct2  # noqa: E402
from ctranslate2_b200.converters.synthetic import LlamaConfig, write_llama_model  # noqa: E402


def rel_rms(a, b):
    return float(np.sqrt(((a - b) ** 2).mean() / (b ** 2).mean()))


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # host-side rendezvous only: the data path is peer memory
    out = {}
    base = os.path.join(tempfile.gettempdir(), "ct2b200_tp_models")
    for quant, ctype, tol in (("int8_float16", "int8_float16", 6e-2), ("float16", "float16", 2e-2),
                             ("awq_gemm", "float16", 2e-2)):
        d = os.path.join(base, quant)
        if rank == 0 and not os.path.exists(os.path.join(d, ".complete")):
            cfg = LlamaConfig(num_layers=3, num_heads=8, num_heads_kv=4, head_dim=128, ffn_dim=2048, vocab_size=2000)
            write_llama_model(d, cfg, quant, seed=21, init_std=0.04)
            open(os.path.join(d, ".complete"), "w").write("ok")
        dist.barrier()
        prompts = np.random.default_rng(5).integers(3, 2000, size=(3, 40))
        single = ct2.Generator(d, device_index=local, compute_type=ctype, max_batch_size=4, max_length=160)
        ref_logits = single.forward_batch(prompts.tolist())
        ref_tokens = [r.sequences_ids[0] for r in single.generate_batch(prompts.tolist(), max_length=24, min_length=24,
                                                                       end_token=[2])]
        single.close()
        g = ct2.Generator(d, device_index=local, compute_type=ctype, max_batch_size=4, max_length=160,
                          tensor_parallel=True)
        logits = g.forward_batch(prompts.tolist())
        tokens = [r.sequences_ids[0] for r in g.generate_batch(prompts.tolist(), max_length=24, min_length=24,
                                                               end_token=[2])]
        tokens2 = [r.sequences_ids[0] for r in g.generate_batch(prompts.tolist(), max_length=24, min_length=24,
                                                                end_token=[2])]
        err = rel_rms(logits, ref_logits)
        gathered = [None] * world
        dist.all_gather_object(gathered, (tokens, float(np.abs(logits).sum())))
        same_across_ranks = all(x == gathered[0] for x in gathered)
        first_tok_match = sum(int(a[0] == b[0]) for a, b in zip(tokens, ref_tokens))
        out[quant] = dict(rel_rms=err, tol=tol, same_across_ranks=same_across_ranks, repeat_identical=tokens == tokens2,
                          first_tokens_equal=first_tok_match, tokens_equal=int(tokens == ref_tokens))
        if os.environ.get("TP_WORKER_REPORT_ONLY"):      # debugging aid: print every case instead of stopping at the first
            if rank == 0:
                print("TP_CASE " + json.dumps({quant: out[quant]}), flush=True)
        else:
            assert same_across_ranks, "ranks disagree"
            assert tokens == tokens2, "second call differs"
            assert err <= tol, (quant, err)
        g.close()
        dist.barrier()
    if rank == 0:
        print("TP_RESULT " + json.dumps(out))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
