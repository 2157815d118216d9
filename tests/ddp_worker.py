"""Worker of tests/test_ddp_gpu.py (not collected by pytest).  Run as
    torchrun --nproc-per-node W tests/ddp_worker.py OUT_DIR          (W ranks: data-parallel, B sentences per rank)
    python tests/ddp_worker.py OUT_DIR --single W                    (1 rank, W*B sentences per batch: the reference)
Both drive kbner_b200.trainer.ModelFinetuner.train ITSELF for one epoch = one optimizer step on the same seeded model and
corpus (dropout off so the two are comparable) and save the parameters; with W ranks the step contains the product's
gradient exchange (distributed.GradExchange: pack to bf16 -> NCCL all-reduce -> AdamW reads the reduced buffer)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

B = 4


def main():
    out_dir = sys.argv[1]
    single = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[2] == "--single" else 0
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    from test_api_gpu import SMALL, _models, _sentences
    from kbner_b200.trainer import ListCorpus, ModelFinetuner
    tagger, emb, params, ocfg = _models(SMALL, 13, seed=31)            # same seed on every rank: identical replicas
    tagger.use_word_dropout = 0.0
    n_ranks = single or world
    sents = _sentences(B * n_ranks, 6, 30, seed=77)
    for s in sents:
        for tok in s.tokens:
            tok.add_tag("ner", "S-T0" if tok.text[0] in "abcde" else "O")
    trainer = ModelFinetuner(tagger, corpus=ListCorpus(sents, [], []))
    # two epochs = two optimizer steps: the second replays the captured forward / backward (chunk) graphs and steps rows of
    # the embedding table that carry momentum from the first
    trainer.train(os.path.join(out_dir, "run-%d-%d" % (n_ranks, rank)), learning_rate=1e-3, lr_rate=10.0,
                  mini_batch_size=B if not single else B * single, max_epochs=2, gradient_accumulation_steps=1,
                  shuffle=False, save_final_model=False, train_with_dev=True)
    sd = {k: v.detach().float().cpu().clone() for k, v in tagger.state_dict().items()}
    torch.save(sd, os.path.join(out_dir, "params-%s-rank%d.pt" % ("single" if single else "ddp", rank)))
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
