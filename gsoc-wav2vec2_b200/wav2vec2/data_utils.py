"""Data pipeline on either side of the hot path: the ``batchify`` semantics of the reference's
``CommonDataLoader`` and its LibriSpeech / TIMIT directory readers (src/data_utils.py:30-324), without tf.data.

What the reference does per utterance (data_utils.py:225-232, 322-324): read the sound file, normalise it with
``Wav2Vec2Processor`` (per utterance, over its REAL samples, i.e. before padding), tokenize the transcript; then
``batchify`` (data_utils.py:52-78): truncate speech / labels to ``audio_maxlen`` / ``labels_maxlen`` (``restrict_to_maxlen``),
``padded_batch`` to exactly ``(audio_maxlen, labels_maxlen)`` with pad values ``(audio_pad_id, labels_pad_id)``,
``drop_remainder`` by default.  (``dataset.shuffle(...)`` at data_utils.py:58-59 discards its result, so the reference never
shuffles; ``seed`` is accepted and ignored here for the same reason.)

Two ways to produce a batch:
  * ``CommonDataLoader.batchify``: host arrays in, host (pinned) tensors out - the reference's arithmetic on the host;
  * ``DeviceBatcher``: RAW (un-normalised) utterances are copied once into a pinned staging buffer, moved to the GPU, and the
    per-utterance normalisation runs there (``w2v2_normalize_utterances``: statistics over each utterance's real samples,
    padded tail written as ``audio_pad_id`` = 0) - the order "normalise, then pad" is kept, the host never touches the samples
    arithmetically.  Note one reference subtlety that both paths keep: normalisation happens BEFORE ``restrict_to_maxlen``,
    so an over-long utterance is normalised over all of its samples and then truncated.
TFRecord shards (make_tfrecords.py) are a TensorFlow serialisation format and stay out of scope.
"""
import os
import struct
from dataclasses import dataclass
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from .processor import Wav2Vec2Processor

Sample = Tuple[np.ndarray, Sequence[int]]


def read_wav(path: str) -> Tuple[np.ndarray, int]:
    """RIFF/WAVE PCM16 or float32 reader -> (mono float32 in [-1, 1), sample_rate); what ``tf.audio.decode_wav``
    (data_utils.py:322-323) and ``soundfile.read`` (:216-217) return for the corpora's 16-bit files."""
    with open(path, "rb") as fh:
        data = fh.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise ValueError(f"{path}: not a RIFF/WAVE file")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        tag, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if tag == b"fmt ":
            fmt = struct.unpack("<HHIIHH", body[:16])
        elif tag == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise ValueError(f"{path}: missing fmt / data chunk")
    kind, channels, rate, _, _, bits = fmt
    if kind == 1 and bits == 16:
        x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
    elif kind == 3 and bits == 32:
        x = np.frombuffer(pcm, dtype="<f4").astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported WAVE encoding (format {kind}, {bits} bits)")
    if channels > 1:
        x = x.reshape(-1, channels)[:, 0]
    return x, rate


class CommonDataLoader:
    """data_utils.py:30-94 - same constructor arguments, ``batchify`` / ``restrict_to_maxlen`` / ``_fetch_and_push_files``."""

    def __init__(self, batch_size: int, buffer_size: int, audio_pad_id: Union[int, float], labels_pad_id: int,
                 audio_maxlen: int, labels_maxlen: int):
        self.batch_size, self.buffer_size = batch_size, buffer_size
        self.audio_pad_id, self.labels_pad_id = float(audio_pad_id), labels_pad_id
        self.audio_maxlen, self.labels_maxlen = audio_maxlen, labels_maxlen
        self.processor = Wav2Vec2Processor(is_tokenizer=False)
        self.tokenizer = Wav2Vec2Processor(is_tokenizer=True)

    def restrict_to_maxlen(self, speech, labels):
        """data_utils.py:80-83: must run before padding."""
        return speech[: self.audio_maxlen], labels[: self.labels_maxlen]

    def batchify(self, dataset: Iterable[Sample], seed: Optional[int] = None, drop_remainder: bool = True
                 ) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        """Yields ``(speech [B, audio_maxlen] fp32, labels [B, labels_maxlen] int32)`` host tensors (pinned when CUDA is
        present) from an iterable of already-normalised ``(speech, labels)`` samples (data_utils.py:52-78)."""
        del seed                                    # the reference's shuffle is a no-op (see module docstring)
        pin = torch.cuda.is_available()
        group: List[Sample] = []
        for sample in dataset:
            group.append(sample)
            if len(group) == self.batch_size:
                yield self._pad(group, pin)
                group = []
        if group and not drop_remainder:
            yield self._pad(group, pin)

    def _pad(self, group: List[Sample], pin: bool):
        speech = torch.full((len(group), self.audio_maxlen), self.audio_pad_id, dtype=torch.float32)
        labels = torch.full((len(group), self.labels_maxlen), self.labels_pad_id, dtype=torch.int32)
        for i, (s, l) in enumerate(group):
            s, l = self.restrict_to_maxlen(torch.as_tensor(np.asarray(s), dtype=torch.float32).reshape(-1),
                                           torch.as_tensor(np.asarray(l), dtype=torch.int32).reshape(-1))
            speech[i, : s.numel()] = s
            labels[i, : l.numel()] = l
        return (speech.pin_memory(), labels.pin_memory()) if pin else (speech, labels)

    def _fetch_and_push_files(self, data_dir: str, file_paths: list, file_pattern: str):
        """data_utils.py:85-97: recursive collection of the files ending in ``file_pattern``."""
        for f in os.listdir(data_dir):
            f = os.path.join(data_dir, f)
            if f.endswith(file_pattern):
                file_paths.append(os.path.abspath(f))
            elif os.path.isdir(f):
                self._fetch_and_push_files(f, file_paths, file_pattern)


class DeviceBatcher:
    """``batchify`` with the normalisation on the GPU.  ``__call__(raw_samples)`` yields
    ``(speech [B, audio_maxlen] fp32 CUDA, labels [B, labels_maxlen] int32 CUDA)``: the same values as
    ``processor(speech)`` per utterance followed by ``CommonDataLoader.batchify``, to fp32 rounding."""

    def __init__(self, loader: CommonDataLoader, device="cuda"):
        if loader.audio_pad_id != 0.0:
            raise ValueError("the device normalise kernel writes the padded tail as 0 (the reference's audio_pad_id)")
        self.loader, self.device = loader, torch.device(device)
        B, L = loader.batch_size, loader.audio_maxlen
        self._raw = torch.zeros((B, L), dtype=torch.float32).pin_memory()
        self._len = torch.zeros((B,), dtype=torch.int32).pin_memory()
        self._lab = torch.zeros((B, loader.labels_maxlen), dtype=torch.int32).pin_memory()

    def _emit(self, n: int, tails):
        from . import ops
        ld = self.loader
        raw = self._raw[:n].to(self.device, non_blocking=True)
        lens = self._len[:n].to(self.device, non_blocking=True)
        labels = self._lab[:n].to(self.device, non_blocking=True)
        speech = ops.normalize_utterances(raw, lens, out=raw)
        # over-long utterances: the statistics must cover the samples that restrict_to_maxlen cuts off afterwards
        for i, (mean, rstd) in tails.items():
            speech[i] = (self._raw[i].to(self.device) - mean) * rstd
        torch.cuda.current_stream(self.device).synchronize()     # the pinned staging buffers are reused by the next batch
        return speech, labels

    def __call__(self, dataset: Iterable[Sample], drop_remainder: bool = True):
        ld, n, tails = self.loader, 0, {}
        for s, l in dataset:
            s = np.asarray(s, dtype=np.float32).reshape(-1)
            l = np.asarray(l, dtype=np.int32).reshape(-1)[: ld.labels_maxlen]
            k = min(s.size, ld.audio_maxlen)
            self._raw[n].zero_()
            self._raw[n, :k] = torch.from_numpy(s[:k])
            self._len[n] = k
            if s.size > ld.audio_maxlen:             # rare: statistics over the FULL utterance (normalise, then truncate)
                m = float(s.mean(dtype=np.float64))
                tails[n] = (m, float(1.0 / np.sqrt(s.astype(np.float64).var() + 1e-5)))
            self._lab[n].fill_(ld.labels_pad_id)
            self._lab[n, : l.size] = torch.from_numpy(l)
            n += 1
            if n == ld.batch_size:
                yield self._emit(n, tails)
                n, tails = 0, {}
        if n and not drop_remainder:
            yield self._emit(n, tails)


@dataclass
class LibriSpeechDataLoaderArgs:
    """data_utils.py:99-129 (``from_tfrecords`` is accepted for signature parity and must stay False: no TensorFlow here)."""
    from_tfrecords: bool = False
    tfrecords: Optional[List[str]] = None
    data_dir: str = "../data/LibriSpeech/test-clean"
    batch_size: int = 16
    buffer_size: int = 10000
    audio_maxlen: int = 400000
    audio_pad_id: int = 0
    labels_maxlen: int = 128
    labels_pad_id: int = 0

    def __post_init__(self):
        if self.from_tfrecords:
            raise NotImplementedError("TFRecord shards are a TensorFlow serialisation format; read the corpus directory instead")
        assert self.data_dir is not None, "You must specify `data_dir` when `from_tfrecords=False`."


@dataclass
class TimitDataLoaderArgs:
    """data_utils.py:132-143."""
    data_dir: str = "../data/timit/data/TRAIN"
    batch_size: int = 16
    buffer_size: int = 10000
    audio_maxlen: int = 400000
    audio_pad_id: int = 0
    labels_maxlen: int = 128
    labels_pad_id: int = 0


def _loader_args(args):
    return (args.batch_size, args.buffer_size, args.audio_pad_id, args.labels_pad_id, args.audio_maxlen, args.labels_maxlen)


class LibriSpeechDataLoader(CommonDataLoader):
    """data_utils.py:146-273: ``<id>.flac`` files matched with the ``<id> TRANSCRIPT`` lines of the ``*.txt`` files.
    ``file_ext`` exists because FLAC decoding needs ``soundfile`` (absent here: ``.wav`` copies of the corpus are read instead)."""

    def __init__(self, args: LibriSpeechDataLoaderArgs, required_sample_rate: int = 16000, file_ext: str = ".flac"):
        super().__init__(*_loader_args(args))
        self.data_dir, self.required_sample_rate, self.file_ext = args.data_dir, required_sample_rate, file_ext
        self._num_samples = None

    def __call__(self, seed: Optional[int] = None, drop_remainder: bool = True, device=None):
        if device is not None:
            return DeviceBatcher(self, device)(self._inputs_generator(self._index(), normalize=False), drop_remainder)
        return self.batchify(self._inputs_generator(self._index()), seed=seed, drop_remainder=drop_remainder)

    def _index(self) -> List[Tuple[str, str]]:
        file_paths: List[str] = []
        self._fetch_and_push_files(self.data_dir, file_paths, self.file_ext)
        names = [os.path.basename(p)[: -len(self.file_ext)] for p in file_paths]
        texts = self._fetch_librispeeh_txt()
        pairs = [(p, texts.pop(n, None)) for p, n in zip(file_paths, names)]
        kept = [pr for pr in pairs if pr[1] is not None]
        print(f"DISCARDING {len(pairs) - len(kept)} samples")
        print(f"LOADED {len(kept)} FILES FROM {self.data_dir}")
        self._num_samples = len(kept)
        return kept

    def __len__(self):
        if self._num_samples is None:
            raise NotImplementedError
        return self._num_samples

    def read_sound(self, file_path: str) -> np.ndarray:
        if file_path.lower().endswith(".flac"):
            try:
                import soundfile as sf
            except ImportError as e:
                raise RuntimeError("FLAC decoding needs `soundfile`; convert the corpus to .wav and pass file_ext='.wav'") from e
            audio, rate = sf.read(file_path, dtype="float32")
        else:
            audio, rate = read_wav(file_path)
        if rate != self.required_sample_rate:
            raise ValueError(f"sample rate (={rate}) of your files must be {self.required_sample_rate}")
        return np.asarray(audio, dtype=np.float32)

    def _inputs_generator(self, text_by_filepath, normalize=True):
        for file_path, text in text_by_filepath:
            speech = self.read_sound(file_path)
            if normalize:
                speech = self.processor(speech).numpy()
            yield speech, np.asarray(self.tokenizer(text), dtype=np.int32)

    def _fetch_librispeeh_txt(self) -> dict:
        """data_utils.py:234-262: ``{file_id: transcript}`` from every ``*.txt`` (lines with fewer than 3 fields are skipped,
        exactly like the reference's ``len(s.split()) > 2``)."""
        txt_paths: List[str] = []
        self._fetch_and_push_files(self.data_dir, txt_paths, ".txt")
        out = {}
        for path in txt_paths:
            with open(path, "r") as fh:
                for line in fh.read().split("\n"):
                    parts = line.split()
                    if len(parts) > 2:
                        out[parts[0]] = " ".join(parts[1:])
        return out


class TimitDataLoader(CommonDataLoader):
    """data_utils.py:265-324: ``X.WAV`` + ``X.TXT`` pairs; the transcript is the TXT content after its two sample-index fields."""

    def __init__(self, args: TimitDataLoaderArgs):
        super().__init__(*_loader_args(args))
        self.data_dir, self.wav_ext, self.txt_ext = args.data_dir, ".WAV", ".TXT"

    def _files(self):
        wavs, txts = [], []
        self._fetch_and_push_files(self.data_dir, wavs, self.wav_ext)
        self._fetch_and_push_files(self.data_dir, txts, self.txt_ext)
        files = sorted(set(f[: -len(self.wav_ext)] for f in wavs) & set(f[: -len(self.txt_ext)] for f in txts))
        print(f"found {len(files)} samples in {self.data_dir}")
        return files

    def __call__(self, seed: Optional[int] = None, drop_remainder: bool = True, device=None):
        files = self._files()
        labels = [self._prepare_labels(self.read_timit_txt(f + self.txt_ext)) for f in files]
        if device is not None:
            gen = ((read_wav(f + self.wav_ext)[0], l) for f, l in zip(files, labels))
            return DeviceBatcher(self, device)(gen, drop_remainder)
        gen = ((self.read_sound(f + self.wav_ext), l) for f, l in zip(files, labels))
        return self.batchify(gen, seed=seed, drop_remainder=drop_remainder)

    def _prepare_labels(self, text: str):
        ids = list(self.tokenizer(text))
        return ids + [self.labels_pad_id] * max(0, self.labels_maxlen - len(ids))

    def read_timit_txt(self, file_path: str) -> str:
        with open(file_path, "r") as fh:
            return " ".join(fh.read().split()[2:])

    def read_sound(self, file_path: str) -> np.ndarray:
        return self.processor(read_wav(file_path)[0]).numpy()
