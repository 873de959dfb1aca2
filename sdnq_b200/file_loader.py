"""Checkpoint tensor readers behind `load_sdnq_model` (reference file_loader.py:6-64): same entry points and `method` names.

  safetensors  one file after the other, tensors materialised straight on `device`
  threaded     one reader thread per file (4 at a time): the files of a sharded checkpoint are read concurrently
  streamer     runai_model_streamer, if that package is installed

`key_mapping` is the {regex: replacement} table Transformers models carry as `_checkpoint_conversion_mapping`; the first
pattern that matches a key renames it."""
import concurrent.futures
import re
import threading

import torch


def map_keys(key: str, key_mapping: dict | None) -> str:
    if key_mapping:
        for pattern, replacement in key_mapping.items():
            renamed, hits = re.subn(pattern, replacement, key)
            if hits:
                return renamed
    return key


def load_safetensors(files, state_dict: dict | None = None, key_mapping: dict | None = None, device="cpu", _lock=None) -> dict:
    from safetensors import safe_open
    state_dict = {} if state_dict is None else state_dict
    for path in files:
        with safe_open(path, framework="pt", device=str(device)) as handle:
            for key in handle.keys():      # noqa: SIM118
                tensor = handle.get_tensor(key)
                if _lock is None:
                    state_dict[map_keys(key, key_mapping)] = tensor
                else:
                    with _lock:
                        state_dict[map_keys(key, key_mapping)] = tensor
    return state_dict


def load_threaded(files, state_dict: dict | None = None, key_mapping: dict | None = None, device="cpu", max_workers: int = 4) -> dict:
    state_dict = {} if state_dict is None else state_dict
    lock = threading.Lock()
    with concurrent.futures.ThreadPoolExecutor(max_workers=max_workers) as pool:
        jobs = [pool.submit(load_safetensors, [path], state_dict, key_mapping, device, lock) for path in files]
        for job in concurrent.futures.as_completed(jobs):
            job.result()
    return state_dict


def load_streamer(files, state_dict: dict | None = None, key_mapping: dict | None = None, device="cpu") -> dict:
    try:
        from runai_model_streamer import SafetensorsStreamer
    except ImportError as e:
        raise ImportError("load method 'streamer' needs the runai_model_streamer package") from e
    state_dict = {} if state_dict is None else state_dict
    with SafetensorsStreamer() as streamer:
        streamer.stream_files(list(files))
        for key, tensor in streamer.get_tensors():
            state_dict[map_keys(key, key_mapping)] = tensor.to(device)
    return state_dict


def load_files(files, state_dict: dict | None = None, key_mapping: dict | None = None, device="cpu", method: str | None = None) -> dict:
    """files: the shards of ONE module's checkpoint."""
    files = [files] if isinstance(files, str) else list(files)
    state_dict = {} if state_dict is None else state_dict
    readers = {"safetensors": load_safetensors, "threaded": load_threaded, "streamer": load_streamer}
    method = method or "safetensors"
    if method not in readers:
        raise ValueError(f"Unsupported loading method: {method}")
    return readers[method](files, state_dict=state_dict, key_mapping=key_mapping, device=device)
