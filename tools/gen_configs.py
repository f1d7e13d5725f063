"""Writes video-spike_b200/config/{model,train}/*.yaml: the same YAML schema as the reference's
config/model/*.yaml and config/train/*.yaml (keys read by the hot path: seed, wandb.use, dirs.*,
training.*, model.model_class, data.modalities.<mod>.input, optimizer.*), trimmed of the entries no
code on this path reads.  Reference config files themselves load unchanged (utils/config_utils.py)."""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
CFG = os.path.join(HERE, "..", "video-spike_b200", "config")

# name -> (input_dim, comment)
MODELS = {
    "linear_video": (1966080, "120 frames x 128 x 128"),
    "linear_whisker-video": (1531200, "120 x 110 x 166 whisker ROI"),
    "linear_me-video": (1966200, "120 x 128 x 128 + 120 motion-energy samples"),
    "linear_whisker-of-video": (4036480, "119 x 106 x 160 x 2 optic flow"),
    "linear_whole-of-video": (3899392, "119 x 128 x 128 x 2 optic flow"),
    "linear_me": (120, "120 motion-energy samples"),
    "linear_whisker-of": (120, "120 optic-flow magnitudes"),
    "linear_marker": (240, "120 x 2"),
    "linear_of-ws": (240, "120 x 2"),
    "linear_whisker-of-2d": (240, "120 x 2"),
}

# name -> (ordered modalities with input flag, lr, epochs)
ALL = ["ap", "video", "choice", "block", "wheel-speed", "whisker-motion-energy"]
TRAIN = {
    "linear_video": ([(m, m == "video") for m in ALL], 5e-5, 200),
    "linear_me-video": ([(m, m in ("video", "whisker-motion-energy")) for m in ALL], 5e-5, 200),
    "linear_me": ([(m, m == "whisker-motion-energy") for m in ALL + ["timestamp"]], 5e-5, 200),
    "linear_marker": ([(m, m in ("wheel-speed", "whisker-motion-energy")) for m in ALL], 5e-5, 200),
    "linear_of-ws": ([(m, m in ("wheel-speed", "of")) for m in ALL + ["whisker-of", "of"]], 5e-5, 200),
    "linear_of": ([(m, m == "of") for m in ALL + ["whisker-of", "of"]], 5e-5, 200),
    "linear_whisker-of": ([(m, m == "whisker-of") for m in ALL + ["whisker-of", "of"]], 5e-5, 200),
    "linear_whisker-of-2d": ([("ap", False), ("video", False), ("whisker-of-2d", True)], 5e-5, 200),
    "linear_whisker-of-video": ([("ap", False), ("video", False), ("whisker-of-video", True)], 5e-5, 200),
    "linear_whole-of-video": ([("ap", False), ("video", False), ("whole-of-video", True)], 5e-5, 200),
    "linear_whisker-video": ([("ap", False), ("video", False), ("whisker-video", True)], 5e-6, 300),
    "rrr": ([(m, m in ("whisker-motion-energy", "whisker-of", "whisker-of-2d"))
             for m in ALL + ["whisker-of", "whisker-of-2d", "timestamp"]], 5e-5, 200),
}


def model_yaml(input_dim, note):
    return f"""# Linear video->spike model (model/linear.py).  input_dim / output_dim are placeholders:
# train.py overwrites both from the first training batch, as the reference does.
model_class: Linear
encoder:
  type: LinearEncoder
  input_dim: {input_dim}   # {note}
  hidden_dims: [256, 128]
  output_dim: 64
  layer_num: 2
decoder:
  type: LinearDecoder
  input_dim: 64
  hidden_dims: [128, 256]
  output_dim: 14400   # 100 time bins x neurons
  layer_num: 2
"""


def train_yaml(mods, lr, epochs):
    lines = ["seed: 42", "wandb:", "  use: false", "  entity: null", "  project: ibl-video", "dirs:",
             "  checkpoint_dir: checkpoints", "  log_dir: results", "  data_dir: data/ibl-video", "training:",
             f"  num_epochs: {epochs}", "  train_batch_size: 16", "  test_batch_size: 16", "  num_workers: 1",
             "model:", "  model_class: null", "data:", "  modalities:"]
    for name, is_input in mods:
        lines += [f"    {name}:", f"      input: {'true' if is_input else 'false'}"]
        if name == "video":
            lines += ["      width: 128", "      height: 128"]
    lines += ["optimizer:", "  lr: " + ("5.e-5" if lr == 5e-5 else "5.e-6"), "  wd: 0.01", "  eps: 1.e-8", "  warmup_pct: 0.15",
              "  div_factor: 10", "  scheduler: cosine", ""]
    return "\n".join(lines)


if __name__ == "__main__":
    os.makedirs(os.path.join(CFG, "model"), exist_ok=True)
    os.makedirs(os.path.join(CFG, "train"), exist_ok=True)
    for name, (dim, note) in MODELS.items():
        open(os.path.join(CFG, "model", name + ".yaml"), "w").write(model_yaml(dim, note))
    for name, (mods, lr, ep) in TRAIN.items():
        seen, uniq = set(), []
        for m, f in mods:
            if m not in seen:
                seen.add(m); uniq.append((m, f))
        open(os.path.join(CFG, "train", name + ".yaml"), "w").write(train_yaml(uniq, lr, ep))
    open(os.path.join(CFG, "accelerate", "default.yaml"), "w").write(
        "# accelerate launch contract: single machine, one process per GPU, no mixed precision wrapper\n"
        "compute_environment: LOCAL_MACHINE\ndistributed_type: 'NO'\nmachine_rank: 0\nmain_training_function: main\n"
        "mixed_precision: 'no'\nnum_machines: 1\nnum_processes: 1\nuse_cpu: false\n")
