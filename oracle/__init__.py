"""CPU oracle for the Synthetic-Audio-Detection inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import anything from this package, and only as the checker (or as the
timed CPU baseline) -- never as part of the shipped path.  The product package
(``synthetic-audio-detection_b200`` / alias ``sad_b200``) must not import it.

Contents
--------
``restatement``  plain torch-fp32 / numpy restatement of the reference algorithm
                 (framing, STFT, mel, dB, standardise, resize, ResNet-18 heads,
                 merge, decision, clip aggregate).  Every function cites the
                 reference file:line it follows.  Travels to the GPU box.
``timm_shim``    a stand-in for the un-installed ``timm`` dependency so that the
                 UNMODIFIED reference modules import (container only).
``reference_api``imports ``/root/reference/modular/source/{inference_runner,
                 model_merger}.py`` unmodified through the shim (container only;
                 ``/root/reference`` does not exist on the GPU box).
``reference_loop`` the reference's own per-clip loop (IR:276-298) on in-memory segments, calling INTO the
                 reference modules (bench.py's ``cpu_baseline`` / ``--impl reference`` / library bar).
``build_ref``    byte-compiles the reference's two hot-path modules into ``oracle/_ref/*.pyc.bin`` (git-ignored
                 binaries that travel to the GPU box; no reference source enters the repository).
``fixtures``     seeded synthetic audio (continuous SURVEY 8d corpus and the class-structured decision
                 corpus) and seeded random-init merged checkpoints (v1, and the v2 decision fixture).
``make_decision_fixture`` calibrates / fits the v2 fixture and writes ``tests/golden/decision_fixture.npz``.
``make_golden``  script that ran the reference here and wrote ``tests/golden/`` (``decisions`` / ``deep`` /
                 ``frontend256`` sub-commands for the long-running round-2 goldens).
``bf16_emulation`` CPU model of the device's rounding points (bf16 or fp16 storage); a debugging aid.

Parity pinning: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), so the restatement is pinned against outputs of the
reference itself executed in the build container (``make_golden.py``), committed
as ``tests/golden/*.npz``.
"""
