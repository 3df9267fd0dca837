#!/usr/bin/env python3
"""Drop-in for the reference's ``modular/source/model_merger.py`` (MM:<line>): merge N two-logit sub-model
checkpoints listed in a CSV into one ``{'state_dict', 'metadata': {'class_names'}}`` file (MM:154-159).

Pure host logic (checkpoint I/O); the optional smoke forward (MM:149-151) runs on the CUDA kernels when a device is
present and is skipped otherwise (there is no CPU fallback)."""
from __future__ import annotations

import argparse
import collections
import copy
import csv
import os
import sys

import torch
import torch.nn as nn

if __package__ in (None, ""):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import sad_b200  # noqa: F401
    from sad_b200.inference_runner import BinaryClassifier, ModularMultiHeadClassifier
else:
    from .inference_runner import BinaryClassifier, ModularMultiHeadClassifier

__all__ = ["BinaryClassifier", "ModularMultiHeadClassifier", "force_separate_parameters", "load_sub_model", "main"]


def force_separate_parameters(model: nn.Module):
    """MM:42-44: give every parameter its own storage."""
    for _, param in model.named_parameters():
        param.data = param.data.clone()


def load_sub_model(checkpoint_path, device, model_name="resnet18"):
    """MM:46-59: load a 2-output model; keys that do not match are silently ignored (strict=False)."""
    model = BinaryClassifier(model_name=model_name)
    ck = torch.load(checkpoint_path, map_location="cpu")
    sd_in = ck["state_dict"]
    model.load_state_dict(sd_in, strict=False)
    model.eval()
    force_separate_parameters(model)
    model._engine = None                      # engines are per-instance device state, never copied
    return copy.deepcopy(model)


def merged_real_class(real_names):
    """MM:137-143: the common value, else the most common one (with a warning)."""
    if len(set(real_names)) == 1:
        return real_names[0]
    merged = collections.Counter(real_names).most_common(1)[0][0]
    print("Warning: Not all real_class values match in CSV; using the most common value:", merged)
    return merged


def read_csv(csv_file):
    """Rows with keys model_filename, synthetic_class, real_class (modular/model-merge-example.csv)."""
    with open(csv_file, newline="") as f:
        return list(csv.DictReader(f))


def main(argv=None):
    parser = argparse.ArgumentParser(description="Merge sub-models into a multi-head classifier with a merged Real output.")
    parser.add_argument("--submodels-folder", type=str, required=True, help="Folder containing sub-model .pth files.")
    parser.add_argument("--csv-file", type=str, required=True,
                        help='CSV file with columns "model_filename", "synthetic_class", and "real_class".')
    parser.add_argument("--model-name", type=str, default="resnet18")
    parser.add_argument("--output-path", type=str, required=True)
    args = parser.parse_args(argv)

    device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    entries = read_csv(args.csv_file)
    if not entries:
        print("No submodels found in CSV file!")
        return

    sub_models, synthetic_names, real_names = [], [], []
    for i, entry in enumerate(entries, start=1):
        model_path = os.path.join(args.submodels_folder, entry["model_filename"])
        print(f"Loading sub-model {i} from {model_path} with synthetic class '{entry['synthetic_class']}' "
              f"and real class '{entry['real_class']}'")
        sub_models.append(load_sub_model(model_path, device, model_name=args.model_name))
        synthetic_names.append(entry["synthetic_class"])
        real_names.append(entry["real_class"])

    merged = ModularMultiHeadClassifier(sub_models).eval()
    final_class_names = synthetic_names + [merged_real_class(real_names)]

    if device.type == "cuda":
        out = merged(torch.randn(2, 3, 512, 512, device=device))                # MM:149-151
        print("Merged model output shape:", out.shape)
    else:
        print("Merged model built (no CUDA device: smoke forward skipped, there is no CPU fallback)")

    torch.save({"state_dict": merged.state_dict(), "metadata": {"class_names": final_class_names}}, args.output_path)
    print(f"Saved merged model with metadata => {args.output_path}")


if __name__ == "__main__":
    main()
