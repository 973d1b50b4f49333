"""Data pipeline around the path: batchify semantics of src/data_utils.py:52-83 and the corpus readers."""
import os
import struct
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gsoc-wav2vec2_b200"))
sys.path.insert(0, ROOT)

from wav2vec2 import Wav2Vec2Processor  # noqa: E402
from wav2vec2.data_utils import (CommonDataLoader, DeviceBatcher, LibriSpeechDataLoader, LibriSpeechDataLoaderArgs,  # noqa: E402
                                 TimitDataLoader, TimitDataLoaderArgs, read_wav)


def write_wav(path, x, rate=16000):
    pcm = np.clip(np.round(x * 32768.0), -32768, 32767).astype("<i2").tobytes()
    with open(path, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(pcm)) + b"WAVE")
        fh.write(b"fmt " + struct.pack("<IHHIIHH", 16, 1, 1, rate, rate * 2, 2, 16))
        fh.write(b"data" + struct.pack("<I", len(pcm)) + pcm)


def test_read_wav_matches_the_reference_fixture():
    """tests/golden/sample.wav is the reference's data/sample.wav; the oracle's reader is the checker."""
    from oracle import w2v2_oracle as O
    x, rate = read_wav(os.path.join(ROOT, "tests", "golden", "sample.wav"))
    assert rate == 16000
    np.testing.assert_array_equal(x, O.read_wav_s16(os.path.join(ROOT, "tests", "golden", "sample.wav")))


def test_batchify_truncates_pads_and_drops_the_remainder():
    ld = CommonDataLoader(batch_size=2, buffer_size=10, audio_pad_id=0, labels_pad_id=0, audio_maxlen=8, labels_maxlen=4)
    data = [(np.arange(1, 6, dtype=np.float32), [5, 6]), (np.arange(1, 13, dtype=np.float32), [7, 8, 9, 10, 11]),
            (np.ones(3, dtype=np.float32), [4])]
    batches = list(ld.batchify(data))
    assert len(batches) == 1                                    # drop_remainder=True (data_utils.py:55)
    speech, labels = batches[0]
    assert speech.shape == (2, 8) and labels.shape == (2, 4) and labels.dtype == torch.int32
    assert speech[0].tolist() == [1, 2, 3, 4, 5, 0, 0, 0]       # padded with audio_pad_id
    assert speech[1].tolist() == [1, 2, 3, 4, 5, 6, 7, 8]       # restrict_to_maxlen
    assert labels.tolist() == [[5, 6, 0, 0], [7, 8, 9, 10]]
    assert len(list(ld.batchify(data, drop_remainder=False))) == 2


def _make_librispeech(root, rng):
    d = root / "19" / "198"
    d.mkdir(parents=True)
    utts = {"19-198-0000": "NORTHANGER ABBEY", "19-198-0001": "THIS LITTLE WORK WAS FINISHED", "19-198-0002": "X"}
    waves = {}
    for i, (uid, _) in enumerate(utts.items()):
        waves[uid] = (0.1 * rng.standard_normal(3000 + 700 * i)).astype(np.float32)
        write_wav(str(d / f"{uid}.wav"), waves[uid])
    (d / "19-198.trans.txt").write_text("\n".join(f"{k} {v}" for k, v in utts.items()) + "\n")
    return utts, waves


def test_librispeech_loader_pairs_audio_with_transcripts(tmp_path):
    rng = np.random.default_rng(0)
    utts, waves = _make_librispeech(tmp_path, rng)
    args = LibriSpeechDataLoaderArgs(data_dir=str(tmp_path), batch_size=2, audio_maxlen=4000, labels_maxlen=40)
    ld = LibriSpeechDataLoader(args, file_ext=".wav")
    (speech, labels), = list(ld())
    assert len(ld) == 2                                          # "X" has fewer than 3 fields: skipped like data_utils.py:255-259
    tok, proc = Wav2Vec2Processor(is_tokenizer=True), Wav2Vec2Processor(is_tokenizer=False)
    got = {}
    for row, lab in zip(speech, labels):
        text = tok.decode([int(t) for t in lab if t != 0], group_tokens=False)
        got[text] = row
    assert sorted(got) == sorted(v for v in utts.values() if len(v.split()) > 1)
    for uid, text in utts.items():
        if text not in got:
            continue
        want = proc(read_wav(str(tmp_path / "19" / "198" / f"{uid}.wav"))[0])[:4000]
        np.testing.assert_allclose(got[text][: want.numel()].numpy(), want.numpy(), atol=1e-6)
        assert float(got[text][want.numel():].abs().sum()) == 0.0
    with pytest.raises(NotImplementedError):
        LibriSpeechDataLoaderArgs(from_tfrecords=True)


def test_timit_loader_reads_wav_txt_pairs(tmp_path):
    rng = np.random.default_rng(1)
    d = tmp_path / "DR1" / "FCJF0"
    d.mkdir(parents=True)
    for name, text in (("SA1", "0 46797 She had your dark suit in greasy wash water all year."), ("SI648", "0 1 A sailboat may have a bone")):
        write_wav(str(d / f"{name}.WAV"), (0.05 * rng.standard_normal(2500)).astype(np.float32))
        (d / f"{name}.TXT").write_text(text + "\n")
    write_wav(str(d / "ORPHAN.WAV"), np.zeros(100, dtype=np.float32))          # no transcript: ignored
    ld = TimitDataLoader(TimitDataLoaderArgs(data_dir=str(tmp_path), batch_size=2, audio_maxlen=3000, labels_maxlen=64))
    (speech, labels), = list(ld())
    assert speech.shape == (2, 3000) and labels.shape == (2, 64)
    tok = Wav2Vec2Processor(is_tokenizer=True)
    texts = sorted(tok.decode(l.tolist(), group_tokens=False) for l in labels)
    assert texts == sorted(["SHE HAD YOUR DARK SUIT IN GREASY WASH WATER ALL YEAR", "A SAILBOAT MAY HAVE A BONE"])
    assert abs(float(speech[0, :2500].mean())) < 1e-5 and float(speech[0, 2500:].abs().max()) == 0.0


@pytest.mark.gpu
def test_device_batcher_equals_processor_then_batchify():
    """Normalise-on-GPU batchify == processor(speech) per utterance + CommonDataLoader.batchify (normalise BEFORE padding,
    pad value 0.0, labels padded to labels_maxlen; over-long utterances normalised over ALL their samples, then cut)."""
    rng = np.random.default_rng(2)
    ld = CommonDataLoader(batch_size=3, buffer_size=10, audio_pad_id=0, labels_pad_id=0, audio_maxlen=5000, labels_maxlen=8)
    raw = [((0.3 * rng.standard_normal(n) + 0.05).astype(np.float32), list(rng.integers(1, 30, size=k)))
           for n, k in ((5000, 3), (1234, 8), (7000, 12), (4999, 1), (16, 2), (5000, 5))]
    proc = Wav2Vec2Processor(is_tokenizer=False)
    want = list(ld.batchify([(proc(s).numpy(), l) for s, l in raw]))
    got = list(DeviceBatcher(ld, "cuda")(raw))
    assert len(got) == len(want) == 2
    for (gs, gl), (ws, wl) in zip(got, want):
        assert gs.is_cuda and gl.is_cuda
        torch.testing.assert_close(gs.cpu(), ws, atol=2e-5, rtol=0)
        assert torch.equal(gl.cpu(), wl)
