"""On-GPU multi-rank parity (sharded result == single-GPU result, bit for bit per utterance).
Run under torchrun with N ranks: every rank computes the WHOLE batch on its own GPU and its shard as a tower would,
and compares loss, gradient, greedy and beam hypotheses of the shard with the rows of the whole-batch result; the
reduced scalars are compared with the whole batch's.  Prints one line per rank."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_batch  # noqa: E402
from neuralasr_b200 import towers  # noqa: E402
from neuralasr_b200.networks import common  # noqa: E402
from neuralasr_b200.utils import sparse_tuple_from  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = 8 * world
g = make_batch(2024, T=400, B=B, C=38, Lmax=80, mode="ragged")
offs = g["label_offsets"]
lens = np.diff(offs)
dense = np.zeros((B, max(int(lens.max()), 1)), np.int32)
for b in range(B):
    dense[b, : lens[b]] = g["label_values"][offs[b]:offs[b + 1]]
labels = sparse_tuple_from(dense, lens)
x = torch.from_numpy(g["logits"]).to(dev)
seq = g["seq_len"]
# whole batch on this GPU
loss_w, grad_w, st_w = common.ctc_loss_and_grad(x, labels, seq)
dec_w, _ = common.decoding(x, seq)
beam_w, lp_w = common.create_model_beam(x, seq)
ler_w = common.label_error_rate(dec_w, labels)
sums_w = common.batch_sums(loss_b=loss_w, ler=ler_w.per_utterance, dist=ler_w.distances)
# this rank's tower
lab_r, seq_r, x_r = towers.shard_inputs(labels, seq, rank, world, logits=g["logits"])
xr = torch.from_numpy(x_r).to(dev)
loss_r, grad_r, st_r = common.ctc_loss_and_grad(xr, lab_r, seq_r)
dec_r, _ = common.decoding(xr, seq_r)
beam_r, lp_r = common.create_model_beam(xr, seq_r)
ler_r = common.label_error_rate(dec_r, lab_r)
sums_r = common.batch_sums(loss_b=loss_r, ler=ler_r.per_utterance, dist=ler_r.distances)
towers.all_reduce_sums(sums_r)
torch.cuda.synchronize()
lo, hi = towers.shard_range(B, rank, world)
ok_loss = torch.equal(loss_r, loss_w[lo:hi])
ok_grad = torch.equal(grad_r, grad_w[:, lo:hi, :])
ok_hyp = torch.equal(dec_r.hyp_len, dec_w.hyp_len[lo:hi]) and torch.equal(dec_r.hyp[:, : dec_r.hyp.shape[1]], dec_w.hyp[lo:hi, : dec_r.hyp.shape[1]])
ok_beam = torch.equal(beam_r.hyp_len, beam_w.hyp_len[lo:hi]) and torch.equal(lp_r, lp_w[lo:hi]) and torch.equal(beam_r.hyp, beam_w.hyp[lo:hi])
ok_dist = torch.equal(ler_r.distances, ler_w.distances[lo:hi])
sw, sr = sums_w.cpu().numpy(), sums_r.cpu().numpy()
ok_sums = bool(np.allclose(sw, sr, rtol=1e-12, atol=0))
print("rank %d/%d rows [%d,%d): loss %s grad %s greedy %s beam %s dist %s | reduced sums %s (whole %s, reduced %s)" % (
    rank, world, lo, hi, ok_loss, ok_grad, ok_hyp, ok_beam, ok_dist, ok_sums, np.array2string(sw, precision=6),
    np.array2string(sr, precision=6)), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if (ok_loss and ok_grad and ok_hyp and ok_beam and ok_dist and ok_sums) else 1)
