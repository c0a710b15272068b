"""Trains a REAL sentencepiece unigram vocabulary (subword pieces, punctuation, digits — the kind of model Flan-T5's spiece.model is)
on English prose found in this image, and writes its (piece, score) list as a fixture for the prompt-assembly tests.

    python tests/golden/make_spm_vocab.py

Why: the token-level prompt assembler (llmrankers/_prompts.py) rests on T5's tokenizer segmenting every whitespace-delimited word on
its own. The synthetic whole-word tokenizer of b200rank.synthetic cannot exercise that claim for subword segmentation; the real
Flan-T5 vocabulary is not available offline. A unigram model trained with the same algorithm and normaliser family is
the closest stand-in: tests/test_host_logic.py::test_prompt_assembler_with_a_trained_subword_vocabulary builds
`transformers.T5Tokenizer(vocab=...)` from this fixture and holds the assembler to whole-string tokenisation on natural text.

Corpus: docstrings of the Python standard library and of numpy (deterministic for a given image; the fixture is committed, so tests do
not depend on it). Output: spm_unigram_vocab.json = {"pieces": [[piece, score], ...], "sample_sentences": [...]}.
"""
import importlib
import io
import json
import os
import pkgutil
import pydoc
import random
import re
import tempfile

import sentencepiece as spm

HERE = os.path.dirname(os.path.abspath(__file__))
MODULES = ["os", "sys", "re", "json", "csv", "math", "random", "itertools", "functools", "collections", "datetime", "pathlib", "subprocess",
           "threading", "asyncio", "logging", "argparse", "unittest", "typing", "socket", "ssl", "http.client", "urllib.request", "email.message",
           "sqlite3", "decimal", "fractions", "statistics", "heapq", "bisect", "textwrap", "string", "struct", "codecs", "io", "shutil", "tempfile",
           "zipfile", "tarfile", "gzip", "hashlib", "hmac", "secrets", "time", "calendar", "locale", "gettext", "inspect", "ast", "dis", "pickle",
           "copy", "enum", "dataclasses", "contextlib", "abc", "numbers", "operator", "weakref", "queue", "multiprocessing", "concurrent.futures",
           "xml.etree.ElementTree", "html.parser", "configparser", "getpass", "platform", "signal", "select", "selectors", "uuid", "ipaddress",
           "numpy", "numpy.linalg", "numpy.random", "numpy.fft", "numpy.ma", "numpy.polynomial"]


def corpus():
    lines = []
    for name in MODULES:
        try:
            text = pydoc.render_doc(importlib.import_module(name), renderer=pydoc.plaintext)
        except Exception:  # noqa: BLE001
            continue
        for para in re.split(r"\n\s*\n", text):
            para = " ".join(para.split())
            if len(para) > 40 and sum(c.isalpha() for c in para) > 0.6 * len(para):
                lines.append(para)
    return lines


def main():
    random.seed(0)
    lines = corpus()
    print(len(lines), "paragraphs,", sum(map(len, lines)) // 1024, "KiB")
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "corpus.txt")
        with open(src, "w") as f:
            f.write("\n".join(lines))
        # T5's spiece.model: unigram, pad 0 / eos 1 / unk 2, no bos, nmt_nfkc normalisation, dummy prefix, whitespace pieces only as prefixes
        spm.SentencePieceTrainer.train(input=src, model_prefix=os.path.join(tmp, "m"), vocab_size=6000, model_type="unigram", pad_id=0, eos_id=1,
                                       unk_id=2, bos_id=-1, character_coverage=0.9995, normalization_rule_name="nmt_nfkc", input_sentence_size=200000,
                                       shuffle_input_sentence=False, num_threads=1)
        sp = spm.SentencePieceProcessor(model_file=os.path.join(tmp, "m.model"))
    pieces = [[sp.id_to_piece(i), float(sp.get_score(i))] for i in range(sp.get_piece_size())]
    assert pieces[0][0] == "<pad>" and pieces[1][0] == "</s>" and pieces[2][0] == "<unk>"
    pieces += [[f"<extra_id_{i}>", 0.0] for i in range(99, -1, -1)]
    sample = random.sample([l for l in lines if 60 < len(l) < 600], 120)
    with open(os.path.join(HERE, "spm_unigram_vocab.json"), "w") as f:
        json.dump({"sentencepiece": spm.__version__, "pieces": pieces, "sample_sentences": sample}, f, ensure_ascii=False)
    print(len(pieces), "pieces;", os.path.getsize(os.path.join(HERE, "spm_unigram_vocab.json")) // 1024, "KiB fixture")


if __name__ == "__main__":
    main()
