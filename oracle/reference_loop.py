"""The reference's own per-clip loop (inference_runner.py:276-298) driven on in-memory segments (TEST INFRASTRUCTURE:
bench.py's CPU / library baselines).  Every arithmetic call below is a call INTO the reference module ``IR`` (or into
the model it built); only ``torchaudio.load`` -- unusable offline -- is bypassed by starting from tensors."""
import os
import tempfile

import torch


def build_model(IR, sd, class_names, device):
    """IR.load_merged_model (IR:77-123) on a checkpoint in the layout model_merger.py:154-159 writes."""
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "merged.pth")
        torch.save({"state_dict": sd, "metadata": {"class_names": list(class_names)}}, path)
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model, meta = IR.load_merged_model(path, device)
    return model, meta


def clip_pass(IR, model, chunks, device, names, threshold=0.5, batch_size=128):
    """IR:276-298: per-segment waveform_to_spectrogram on the CPU, one H2D of all windows, the model in mini-batches
    of 128, then interpret_multihead_logits row by row.  Returns (labels, raw_probs)."""
    spec_cfg = IR.SpectrogramConfig(2048, 512, 128, 20, 12000, 80, "slaney")                       # IR:259
    specs = [IR.waveform_to_spectrogram(c, 32000, spec_cfg) for c in chunks]                       # IR:276-279
    all_specs = torch.cat(specs, dim=0).to(device)                                                 # IR:280
    outputs = []
    with torch.no_grad():
        for start in range(0, all_specs.size(0), batch_size):                                      # IR:284-288
            outputs.append(model(all_specs[start:start + batch_size]))
    outputs = torch.cat(outputs, dim=0)
    labels, raw = [], []
    for row in outputs:                                                                            # IR:294-298
        lab, probs = IR.interpret_multihead_logits(row, threshold, names[:-1], names[-1])
        labels.append(lab)
        raw.append(probs)
    return labels, raw
