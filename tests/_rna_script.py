"""The RNA training script as a call sequence, for the whole-script parity test.

`run_like_script(models, ...)` issues, in the script's own order, exactly the calls that the UNMODIFIED reference script
    /root/reference/2_GeneExpression/1_GeneExpress_train.py
makes into the modules this repository replaces (`from models import cox_loss, RNAOnlyModel`, :40): seeding (:227-228),
model construction (:247-257), datasets / RandomSampler / DataLoader (:266-283), `.to(device)` (:288), the two Adam
parameter groups (:303-305), `train_model` (:126-222: train step, then `evaluate` on train and val every epoch, val /
val / test at the end) and `evaluate` (:52-99: eval-mode forward + cox_loss + the per-case grouping of get_survival_CI,
:101-124).  It returns what the script hands to `lifelines.utils.concordance_index` at every evaluate call:
(survival_months, -score, vital_status) per case id in sorted order.

Why a mirror and not the script itself on the GPU: the reference tree cannot travel to the GPU box (and must not be
copied into this repository), and the drop-in has no CPU path to run the script here.  The mirror is pinned instead:
`tools/make_golden.py rna_script` runs the unmodified script on CPU (third-party imports stubbed, the concordance stub
recording its arguments), runs this mirror with the reference's own `models` module on the same inputs, and asserts that
both recordings are IDENTICAL before writing tests/golden/rna_script_reference.npz.

Dropout: the script trains with nn.Dropout() and Adam's first steps move every weight by +-lr whatever the gradient's
size, so after one step on 24 samples the scores depend on the dropout masks at the 10 % level (measured: 9.2 % between
torch's CPU stream and the kernels' Philox stream) - a run with dropout cannot be compared tighter than that with ANY
other random stream.  The golden therefore also holds the mirror's recording with dropout_p = 0 (same initial weights,
same sampler order, reference `models` module, CPU): that variant is deterministic and is compared at 1e-2.
"""
import os

import numpy as np
import torch
import torch.nn as nn
from torch.optim import Adam
from torch.utils.data import RandomSampler

N_GENES = 12778
SPLITS = {"train": 24, "val": 12, "test": 12}
CONFIG = {"batch_size": 128, "num_workers": 0, "num_epochs": 2, "lr_rna": 1e-05, "lr_mlp": 1e-05, "weight_decay": 1e-05,
          "flag": "rna_model", "restore_path": "", "model_path": ""}   # ExampleConfigs/config_rna_train.json values
SEED = 3333                                                            # the script's default --seed (:335)


def synthetic_split(name):
    """(cases, survival_months, vital_status, rna) of one synthetic split - seeded, identical on every box."""
    n = SPLITS[name]
    rng = np.random.default_rng({"train": 101, "val": 202, "test": 303}[name])
    cases = [f"TCGA-{name[:2].upper()}-{i:04d}" for i in range(n)]
    months = np.round(rng.uniform(1.0, 120.0, n), 1).astype(np.float32)
    vital = (rng.uniform(size=n) < 0.6).astype(np.float32)
    rna = np.round(rng.standard_normal((n, N_GENES)), 4).astype(np.float32)   # log + z-score like values
    return cases, months, vital, rna


def write_csvs(directory):
    """The three CSVs in the layout of ExampleData/rna_example.csv (case, survival_months, vital_status, grade_binary,
    rna_0 ...).  Returns {split: path}."""
    import pandas as pd
    paths = {}
    for name in SPLITS:
        cases, months, vital, rna = synthetic_split(name)
        df = pd.DataFrame({"case": cases, "survival_months": months, "vital_status": vital.astype(np.int64),
                           "grade_binary": np.zeros(len(cases), np.int64)})
        df = pd.concat([df, pd.DataFrame(rna, columns=[f"rna_{i}" for i in range(N_GENES)])], axis=1)
        paths[name] = os.path.join(directory, f"rna_{name}.csv")
        df.to_csv(paths[name], index=False, float_format="%.4f")
    return paths


class _RNADataset(torch.utils.data.Dataset):
    """What RNADataset (2_GeneExpression/datasets.py:11-58) yields for the CSVs of write_csvs()."""

    def __init__(self, name):
        cases, months, vital, rna = synthetic_split(name)
        self.data = [{"case": c, "survival_months": np.float32(m), "vital_status": np.float32(v), "grade_binary": 0,
                      "rna_data": torch.tensor(r, dtype=torch.float32)} for c, m, v, r in zip(cases, months, vital, rna)]

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        item = self.data[idx].copy()
        item["idx"] = idx
        return item


def _grouped(output_list, ids_list, survival_months, vital_status):
    """get_survival_CI (:101-124) up to the concordance call: its three arguments."""
    ids_unique = sorted(list(set(ids_list)))
    id_to_scores, id_to_months, id_to_vital = {}, {}, {}
    for i in range(len(output_list)):
        k = ids_list[i]
        id_to_scores[k] = id_to_scores.get(k, []) + [output_list[i, 0]]
        id_to_months[k] = survival_months[i]
        id_to_vital[k] = vital_status[i]
    score = np.array([np.mean(id_to_scores[k]) for k in ids_unique])
    return (np.array([id_to_months[k] for k in ids_unique]), -score, np.array([id_to_vital[k] for k in ids_unique]))


def _evaluate(models, model, loader, device, record):
    model.eval()
    outs, months, vitals, cases, losses = [], [], [], [], []
    for batch in loader:
        inputs = batch["rna_data"].to(device)
        sm = batch["survival_months"].to(device).float()
        vs = batch["vital_status"].to(device).float()
        with torch.no_grad():
            outputs = model.forward(inputs)
            loss = models.cox_loss(outputs.view(-1), sm.view(-1), vs.view(-1))
        losses.append(loss.item())
        outs.append(outputs.detach().cpu().numpy())
        months.append(sm.detach().cpu().numpy())
        vitals.append(vs.detach().cpu().numpy())
        cases.append(batch["case"])
    cases = [c for cb in cases for c in cb]
    record.append(_grouped(np.concatenate(outs, axis=0), cases, np.concatenate(months), np.concatenate(vitals)))
    return float(np.mean(losses))


def run_like_script(models, device, optimizer_hook=None, dropout_p=0.5):
    """Returns (recorded concordance arguments per evaluate call, printed TRAIN losses, final state_dict).
    dropout_p = 0.5 is the script (nn.Dropout()); 0.0 is the deterministic variant of the parity test (dropout draws
    nothing from the generator, so the initial weights and the sampler order are those of the script)."""
    np.random.seed(SEED)
    torch.random.manual_seed(SEED)
    model_rna = nn.Sequential(nn.Dropout(dropout_p), nn.Linear(N_GENES, 4096), nn.ReLU(), nn.Dropout(dropout_p),
                              nn.Linear(4096, 2048))
    combine_mlp = nn.Sequential(nn.Linear(2048, 1))
    model = models.RNAOnlyModel(model_rna, combine_mlp)
    datasets = {x: _RNADataset(x) for x in ("train", "val", "test")}
    samplers = {x: RandomSampler(datasets[x]) for x in ("train", "val", "test")}
    loaders = {x: torch.utils.data.DataLoader(datasets[x], batch_size=CONFIG["batch_size"], sampler=samplers[x],
                                              num_workers=CONFIG["num_workers"]) for x in ("train", "val", "test")}
    model = model.to(device)
    optimizer = Adam([{"params": [p for p in model_rna.parameters() if p.requires_grad], "lr": CONFIG["lr_rna"]},
                      {"params": [p for p in combine_mlp.parameters() if p.requires_grad], "lr": CONFIG["lr_mlp"]}],
                     weight_decay=CONFIG["weight_decay"])
    if optimizer_hook is not None:
        optimizer = optimizer_hook(optimizer)
    record, train_losses = [], []
    best_val, best_state = np.inf, None
    for epoch in range(CONFIG["num_epochs"]):
        model.train()
        running, seen = 0.0, 0.0
        for batch in loaders["train"]:
            inputs = batch["rna_data"].to(device)
            sm = batch["survival_months"].to(device).float()
            vs = batch["vital_status"].to(device).float()
            optimizer.zero_grad()
            outputs = model(inputs)
            loss = models.cox_loss(outputs.view(-1), sm.view(-1), vs.view(-1))
            loss.backward()
            optimizer.step()
            vsum = vs.sum().item()
            running += loss.item() * vsum
            seen += vsum
        train_losses.append(running / seen)
        _evaluate(models, model, loaders["train"], device, record)
        val_loss = _evaluate(models, model, loaders["val"], device, record)
        if val_loss < best_val:
            best_val = val_loss
            best_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
    _evaluate(models, model, loaders["val"], device, record)
    last_state = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    model.load_state_dict(best_state)
    _evaluate(models, model, loaders["val"], device, record)
    _evaluate(models, model, loaders["test"], device, record)
    return record, train_losses, last_state
