"""Host-side logic of the streaming call (no GPU): prompt template, sentinel tokenisation and stop rule
against fixtures produced by the reference's own code; dialogue expansion / KV prefix planning; and the
world_size-2 metric exchange over gloo."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from streammind_b200 import dist_util
from streammind_b200.constants import MMODAL_TOKEN_INDEX
from streammind_b200.conversation import conv_templates
from streammind_b200.mm_utils import KeywordsStoppingCriteria, tokenizer_MMODAL_token
from streammind_b200.model import DialogueCache, expand_dialogue

G = os.path.join(os.path.dirname(__file__), "golden")


class ToyTokenizer:
    pad_token_id, bos_token_id, eos_token_id = 0, 1, 2

    def __init__(self, vocab):
        self.vocab = dict(vocab)

    def __call__(self, text):
        ids = [self.bos_token_id]
        for w in text.replace("\n", " \n ").split(" "):
            if w:
                ids.append(self.vocab.setdefault(w, len(self.vocab) + 3))
        return type("Enc", (), {"input_ids": ids})()

    def batch_decode(self, ids, skip_special_tokens=True):
        inv = {v: k for k, v in self.vocab.items()}
        return [" ".join(inv.get(int(t), "?") for t in row if not (skip_special_tokens and int(t) in (0, 1, 2)))
                for row in ids]


@pytest.fixture(scope="module")
def fx():
    return json.load(open(os.path.join(G, "host_logic.json")))


def test_prompt_template_matches_reference(fx):
    conv = conv_templates["mistral_instruct"].copy()
    conv.append_message(conv.roles[0], "<video>\n")
    conv.append_message(conv.roles[1], None)
    assert conv.get_prompt() == fx["prompts"][0]


def test_sentinel_tokenisation_matches_reference(fx):
    tok = ToyTokenizer({"</s>": 2})
    for p, ids in zip(fx["prompts"], fx["ids"]):
        assert tokenizer_MMODAL_token(p, tok, MMODAL_TOKEN_INDEX["VIDEO"]) == ids
    t = tokenizer_MMODAL_token(fx["prompts"][0], tok, -201, return_tensors="pt")
    assert t.dtype == torch.long and t.tolist() == fx["ids"][0]
    with pytest.raises(ValueError):
        tokenizer_MMODAL_token(fx["prompts"][0], tok, -201, return_tensors="np")


def test_stop_rule_matches_reference(fx):
    tok = ToyTokenizer(fx["vocab"])
    inp = torch.tensor([fx["ids"][0]])
    sc = KeywordsStoppingCriteria(["</s>"], tok, inp)
    assert [k.tolist() for k in sc.keyword_ids] == fx["keyword_ids"]
    assert sc.single_token_ids == [2] and not sc.needs_host_check
    for case in fx["stop_cases"]:
        assert bool(sc(torch.tensor([fx["ids"][0] + case["tail"]]), None)) == case["stop"]


def test_expand_dialogue_and_prefix_planning():
    V = MMODAL_TOKEN_INDEX["VIDEO"]
    ids1 = [1, 10, 11, V, 12]
    items1 = expand_dialogue(ids1, [3])
    assert items1 == [("t", 1), ("t", 10), ("t", 11), ("f", 0), ("f", 1), ("f", 2), ("t", 12)]
    dc = DialogueCache()
    assert dc.plan(items1) == 0
    dc.commit(items1, [40, 41, 42])                        # 42 was never fed back
    ids2 = ids1 + [40, 41, 42, 2, 13, V, 14]
    items2 = expand_dialogue(ids2, [3, 5])
    assert items2[7:] == [("t", 40), ("t", 41), ("t", 42), ("t", 2), ("t", 13), ("f", 3), ("f", 4), ("t", 14)]
    assert dc.plan(items2) == len(items1) + 2              # dialogue + 40, 41 re-used
    # re-tokenised text differs from the generated ids: shorter prefix
    ids2b = ids1 + [40, 99, 42, 2, 13, V, 14]
    assert dc.plan(expand_dialogue(ids2b, [3, 5])) == len(items1) + 1
    # identical dialogue: one position is always left to prefill
    dc2 = DialogueCache(); dc2.commit(items1, [7])
    assert dc2.plan(items1) == len(items1) - 1
    with pytest.raises(ValueError):
        expand_dialogue(ids2, [3])
    assert expand_dialogue([], []) == []


def test_stream_partition_and_aggregate():
    assert dist_util.stream_ids_for_rank(8, 1, 4) == [1, 5]
    assert sum(len(dist_util.stream_ids_for_rank(8, r, 3)) for r in range(3)) == 8
    per = [{"frames": 64.0, "ms": 100.0}, {"frames": 64.0, "ms": 200.0}]
    assert dist_util.aggregate_throughput(per) == pytest.approx(128 / 0.2)
    assert dist_util.gather_metrics({"a": 1.0}) == [{"a": 1.0}]
    assert dist_util.max_over_ranks(3.0) == 3.0


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    r, w = dist_util.init("gloo")
    dist_util.barrier()
    worst = dist_util.max_over_ranks(10.0 * (rank + 1))
    per = dist_util.gather_metrics({"frames": 64.0, "ms": 10.0 * (rank + 1)})
    q.put((rank, worst, dist_util.aggregate_throughput(per), dist_util.stream_ids_for_rank(5, r, w)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_two_rank_metric_exchange_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=100) for _ in ps)
    [p.join(30) for p in ps]
    assert [r[1] for r in res] == [20.0, 20.0]                         # max over ranks
    assert res[0][2] == res[1][2] == pytest.approx(128 / 0.02)         # whole-job frames / worst time
    assert res[0][3] == [0, 2, 4] and res[1][3] == [1, 3]
