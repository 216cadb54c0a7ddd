"""Laconic oblivious transfer on top of the vector layer — mirror of tests/laconic_ot.rs:15-113."""
from __future__ import annotations

from .kzg import KZGSetup
from .types import G1, Radix2EvaluationDomain
from .vec import PADDING_LEN, vec_commit, vec_decrypt, vec_encrypt


class Receiver:
    """tests/laconic_ot.rs:15-58"""

    def __init__(self, kzg_setup: KZGSetup, rng, choices):
        self.kzg_setup = kzg_setup
        self.choices = [int(c) for c in choices]
        self.commitment, self.proofs = vec_commit(rng, kzg_setup, self.choices)

    def receive(self, encrypted_sets):
        n_choices = len(encrypted_sets[0])
        chosen = [encrypted_sets[0][i] if self.choices[i] == 0 else encrypted_sets[1][i] for i in range(n_choices)]
        return vec_decrypt(self.proofs, chosen, ctx=self.kzg_setup.ctx)


class Sender:
    """tests/laconic_ot.rs:60-113"""

    def __init__(self, kzg_setup: KZGSetup, commitment: G1):
        self.kzg_setup = kzg_setup
        self.commitment = commitment

    def send(self, rng, private_set):
        n_values = len(private_set[0])
        elements = Radix2EvaluationDomain(n_values + PADDING_LEN).elements()
        ct0 = vec_encrypt(rng, self.kzg_setup, self.commitment, elements, [0] * n_values, private_set[0])
        ct1 = vec_encrypt(rng, self.kzg_setup, self.commitment, elements, [1] * n_values, private_set[1])
        return [ct0, ct1]
