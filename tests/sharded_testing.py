"""Test-only pieces for the row-sharded path: an oracle-backed backend (CPU, float64 maths) that lets the routing /
exchange logic of ShardedBPR run under gloo without a GPU, and an in-process multi-rank driver that runs G shards on
ONE device in lockstep (exchanges done by slicing) so the CUDA shard kernels can be checked on a single GPU."""
import numpy as np
import torch


class OracleBackend:
    """Same contract as sharded.CudaBackend, computed with the closed form of oracle/bpr_oracle.py (true-space
    tables, dense decay).  TEST INFRASTRUCTURE ONLY."""

    def __init__(self):
        self._loss = 0.0
        self.loss = torch.zeros(1, dtype=torch.float64)

    def gather_rows(self, Q_local, rows_local):
        return Q_local[rows_local.long()]

    def shard_step(self, P_local, cache, tri_local, lr, wd):
        P = P_local.numpy().astype(np.float64)
        C = cache.numpy().astype(np.float64)
        t = tri_local.numpy().astype(np.int64).reshape(-1, 3)
        u, i, j = t[:, 0], t[:, 1], t[:, 2]
        pu, qi, qj = P[u], C[i], C[j]
        x = (pu * (qi - qj)).sum(-1)
        s = 1.0 / (1.0 + np.exp(x))
        dP = np.zeros_like(P)
        G = np.zeros_like(C)
        np.add.at(dP, u, s[:, None] * (qi - qj))
        np.add.at(G, i, s[:, None] * pu)
        np.add.at(G, j, -s[:, None] * pu)
        P_local.copy_(torch.from_numpy((P * (1 - lr * wd) + lr * dP).astype(np.float32)))
        self._loss += float(np.logaddexp(0.0, -x).sum())
        self.loss[0] = self._loss
        return torch.from_numpy(G.astype(np.float32))

    def owner_apply(self, Q_local, rows_local, grads, lr, wd):
        Q = Q_local.numpy().astype(np.float64)
        dQ = np.zeros_like(Q)
        np.add.at(dQ, rows_local.numpy().astype(np.int64), grads.numpy().astype(np.float64))
        Q_local.copy_(torch.from_numpy((Q * (1 - lr * wd) + lr * dQ).astype(np.float32)))

    def materialize(self, P_local, Q_local):
        pass

    def check(self):
        pass


def route(triples, layout, rank):
    """Triples of the global batch whose user is owned by `rank`, with the user column made local."""
    u0, u1 = layout.user_range(rank)
    m = (triples[:, 0] >= u0) & (triples[:, 0] < u1)
    t = triples[m].copy()
    t[:, 0] -= u0
    return t


def lockstep_step(shards, batches):
    """One sharded step of all `shards` (ranks 0..G-1 living in this process): phases in lockstep, exchanges by
    slicing.  Mirrors ShardedBPR.step + DistComm.exchange exactly."""
    G = len(shards)
    plans = [s.plan(b) for s, b in zip(shards, batches)]                      # (ids, send_counts, tri_local)
    off = [np.concatenate([[0], np.cumsum(p[1])]) for p in plans]

    def gather_for(owner, what):        # rows every rank r sends to `owner`, concatenated in rank order
        return torch.cat([what[r][off[r][owner]:off[r][owner + 1]] for r in range(G)])

    ids = [p[0] for p in plans]
    recv_ids = [gather_for(o, ids) for o in range(G)]
    rows = [shards[o].serve(recv_ids[o]) for o in range(G)]
    # send the rows back: owner o holds blocks in requester order with sizes plans[r].send_counts[o]
    caches = []
    for r in range(G):
        parts = []
        for o in range(G):
            start = sum(plans[q][1][o] for q in range(r))
            parts.append(rows[o][start:start + plans[r][1][o]])
        caches.append(torch.cat(parts))
    grads = [shards[r].compute(plans[r][2], caches[r]) for r in range(G)]   # an empty batch still decays the rank's rows
    grads_in = [gather_for(o, grads) for o in range(G)]
    for o in range(G):
        shards[o].apply(recv_ids[o], grads_in[o])
