"""Sampled decoding (teacher_forcing_prob < 1, `sample()`): vae/model.py:463-472,484-512."""


def run_forward_sampled(model, inputs, lengths, coins, eps=None):
    raise NotImplementedError("sampled decoding (teacher_forcing_prob < 1) is SURVEY.md 8f n1")


def run_sample(model, z, max_length):
    raise NotImplementedError("sample() is SURVEY.md 8f n1")
