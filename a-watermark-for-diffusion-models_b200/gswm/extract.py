"""Drop-in for the codec half of the reference's ``extract.py`` (extract.py:72-110): same names,
arguments and return values; the DDIM-inversion / CLI half (extract.py:31-70, 112-211) stays
upstream and keeps calling these two functions.

    from gswm.extract import recover_exactracted_message, calculate_bit_accuracy
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, codec


def _as_tensor(reversed_latents) -> torch.Tensor:
    if isinstance(reversed_latents, torch.Tensor):
        return reversed_latents.detach()
    return torch.from_numpy(np.ascontiguousarray(reversed_latents))


def _rejection(flag: int):
    """The exception the reference ends in for a latent the kernel flagged (extract.py:83 / :86)."""
    if flag & _lib.FLAG_NAN:
        return ValueError("cannot convert float NaN to integer")
    if flag & _lib.FLAG_RANGE:
        return ValueError("invalid literal for int() with base 2: int(norm.cdf(z) * 2) == 2 for an element >= 8.2924")
    return None


def _decode(reversed_latents, args, batched: bool):
    """Shared front end: dtype / length checks in the reference's order, one extract launch, results + per-latent flags."""
    if int(args.l) != 1:
        # extract.py:84-86 emits digits >= 2 into a base-2 parse for l > 1: non-functional upstream
        raise ValueError("invalid literal for int() with base 2 (window size l must be 1)")
    z = _as_tensor(reversed_latents)
    z = z.reshape(z.shape[0], -1) if batched else z.reshape(1, -1)
    if z.dtype not in (torch.float32, torch.float16, torch.bfloat16, torch.float64):
        z = z.to(torch.float32)                     # integer tensors etc.; float64 is decoded as float64 on the device
    msg_bits = int(args.message_length)
    if msg_bits <= 0 or z.shape[1] % msg_bits:
        # extract.py:94,98: the short last segment is indexed past its end
        raise IndexError("string index out of range")
    if (z.shape[1] * z.element_size()) % 16:
        z = z.to(torch.float32)                     # odd-sized 16-bit rows: the kernel fetches rows in 16-byte pieces (exact upcast)
    km = codec.KeyMaterial.make(args.key, args.nonce, None, msg_bits)
    if z.is_cuda:
        res = codec.extract_batch(z, km)
        return res, res.flags.cpu().numpy()
    # a host tensor (what extract.py:48,70 hands over): the host-buffer pipe -- one pinned staging copy in, the decoded bytes
    # and flags written straight to host memory by the kernel, no torch tensor in between
    from . import _embed_common as common
    with common._pipes_lock:
        msgs, _, _, counters, flags = common.host_pipe(z.shape[1]).extract(z, km)
    res = codec.ExtractResult(torch.from_numpy(msgs), None, None, torch.from_numpy(counters), torch.from_numpy(flags), msg_bits)
    return res, flags


def recover_exactracted_message(reversed_latents, args) -> str:
    """extract.py:72-101.  ``reversed_latents``: the inverted latent, any shape that flattens (C
    order, as np.nditer walks it) to the embedded element order -- normally a (1, 4, h, w) fp16 CPU
    tensor (extract.py:48,70).  ``args`` carries ``key``, ``nonce`` (bytes), ``l`` and
    ``message_length`` (any integer that divides the latent size, as in the reference).  Returns the '0'/'1' string
    of length message_length.

    Raises ValueError where the reference does: NaN input (int(nan)) and any z with
    norm.cdf(z) * 2 rounding to 2 (z >= 8.2924, +inf), whose digit '2' breaks int(..., 2) at
    extract.py:86 -- both found by the extract kernel in the same pass (GSWM_FLAG_*), not by extra scans of the tensor.
    """
    res, flags = _decode(reversed_latents, args, batched=False)
    err = _rejection(int(flags[0]))
    if err is not None:
        raise err
    return res.bit_strings()[0]


def calculate_bit_accuracy(original_message_hex, extracted_message_bin):
    """extract.py:103-110 (host string arithmetic, O(message length))."""
    width = 4 * len(original_message_hex)                       # zero-padded: 4 bits per hex digit
    reference_bits = format(int(original_message_hex, 16), f"0{width}b")
    n = min(len(reference_bits), len(extracted_message_bin))      # silently truncates to the shorter one
    reference_bits = reference_bits[:n]
    agree = sum(a == b for a, b in zip(reference_bits, extracted_message_bin[:n]))
    return reference_bits, agree / n


def write_batch_info(result_file, args):
    """extract.py:166-175: the header block of result.txt."""
    from datetime import datetime

    result_file.write("=" * 40 + "Batch Info" + "=" * 40 + "\n")
    result_file.write(f"Time,{datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
    for field in ("key_hex", "nonce_hex", "original_message_hex", "num_inference_steps", "scheduler"):
        result_file.write(f"{field},{getattr(args, field)}\n")
    result_file.write("=" * 40 + "Batch Start" + "=" * 40 + "\n")


def evaluate_latents(names, reversed_latents, args, result_file=None):
    """Batched form of the per-image loop in extract.process_single_directory (extract.py:134-163): decode every
    inverted latent of ``reversed_latents`` [B, ...] in one launch and report like the reference does --
    ``<name>, Bit Accuracy, <acc>`` per image, then ``Average Bit Accuracy, <mean>``.  An image the reference's
    ``recover_exactracted_message`` would raise on (NaN / z >= 8.2924) gets the reference's
    ``Error processing <name>: <exception>`` line instead, is printed, and is left out of the average, exactly like the
    per-image ``try / except`` at extract.py:148-155.  Returns (extracted bit strings -- None for a rejected image --,
    accuracies -- None likewise --, average over the accepted ones)."""
    res, flags = _decode(reversed_latents, args, batched=True)
    strings = res.bit_strings()
    accs = []
    for i, (name, s) in enumerate(zip(names, strings)):
        err = _rejection(int(flags[i]))
        if err is not None:
            strings[i] = None
            accs.append(None)
            print(f"Error processing {name}: {err}\n")
            if result_file is not None:
                result_file.write(f"Error processing {name}: {err}\n")
            continue
        acc = calculate_bit_accuracy(args.original_message_hex, s)[1]
        accs.append(acc)
        if result_file is not None:
            result_file.write(f"{name}, Bit Accuracy, {acc}\n")
    good = [a for a in accs if a is not None]
    avg = sum(good) / len(good) if good else 0.0
    if result_file is not None and good:
        result_file.write(f"Average Bit Accuracy, {avg}\n\n")
        result_file.write("=" * 40 + "Batch End" + "=" * 40 + "\n")
    return strings, accs, avg
