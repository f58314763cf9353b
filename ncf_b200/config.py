"""Run configuration with the reference's names and defaults (reference src/utils/config.py:20-51,
configs/experiments/neumf.yaml).  Unlike the reference singleton this has no import-time side
effects: the YAML is optional (defaults = the shipped neumf.yaml values) and output directories are
created when something is first written."""
from __future__ import annotations

import os
from pathlib import Path

_DEFAULTS = dict(
    raw_data="data/raw/u.data", train_rating="data/processed/u.train.rating",
    test_rating="data/processed/u.test.rating", test_negative="data/processed/u.test.negative",
    user_num=944, item_num=1683, factor_num=32, num_layers=2, dropout=0.0, model_type="NeuMF-end",
    batch_size=256, epochs=20, lr=0.001, num_ng=4, test_num_ng=99, top_k=10,
    temperature=2.0, alpha=0.5, output_dir="results")

_SECTIONS = {"data": ("raw_data", "train_rating", "test_rating", "test_negative"),
             "model": ("user_num", "item_num", "factor_num", "num_layers", "dropout"),
             "training": ("batch_size", "epochs", "lr", "num_ng", "test_num_ng", "top_k"),
             "distillation": ("temperature", "alpha")}


class Config:
    def __init__(self, config_path=None):
        path = Path(config_path or os.environ.get("NCF_CONFIG", "configs/experiments/neumf.yaml"))
        self.config_path = path
        values = dict(_DEFAULTS)
        if path.exists():
            import yaml
            raw = yaml.safe_load(path.read_text()) or {}
            for section, keys in _SECTIONS.items():
                for k in keys:
                    if k in raw.get(section, {}):
                        values[k] = raw[section][k]
            values["model_type"] = raw.get("model", {}).get("type", values["model_type"])
            values["output_dir"] = raw.get("output", {}).get("dir", values["output_dir"])
        for k, v in values.items():
            setattr(self, k, v)
        for k in ("raw_data", "train_rating", "test_rating", "test_negative", "output_dir"):
            setattr(self, k, Path(getattr(self, k)))
        self.log_dir = self.output_dir / "logs"
        self.model_dir = self.output_dir / "models"
        self.figure_dir = self.output_dir / "figures"

    def ensure_dirs(self):
        for d in (self.output_dir, self.log_dir, self.model_dir, self.figure_dir):
            d.mkdir(parents=True, exist_ok=True)

    def get(self, key, default=None):
        return getattr(self, key, default)


config = Config()
