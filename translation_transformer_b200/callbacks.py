"""`PredictionWriter` with the reference's CSV format (src/callbacks.py:42-64), usable with or
without Lightning: header `source,target,prediction_1..n`, one row per query, service tokens
dropped and decoding stopped at the first EOS."""
from __future__ import annotations

from pathlib import Path

try:  # Lightning is optional in this image; the writer works standalone through `write`
    from pytorch_lightning.callbacks import BasePredictionWriter as _Base
except Exception:  # pragma: no cover - exercised when Lightning is absent
    class _Base:  # type: ignore
        def __init__(self, write_interval="batch"):
            self.interval = write_interval


class PredictionWriter(_Base):
    def __init__(self, output_dir, write_interval="batch"):
        super().__init__(write_interval)
        self.output_path = Path(output_dir).resolve()
        self.output_path.unlink(missing_ok=True)
        self.output_path.parent.mkdir(exist_ok=True)

    def write(self, tokenizer, prediction, batch) -> None:
        prediction_np = prediction.cpu().numpy()
        _, n_predictions, _ = prediction_np.shape
        with open(self.output_path, "a") as f:
            if f.tell() == 0:
                print(",".join(["source", "target"] + [f"prediction_{i}" for i in range(1, n_predictions + 1)]), file=f)
            src = batch["src_tokens"].cpu().numpy()
            tgt = batch["tgt_tokens"].cpu().numpy()
            for i, (s, t) in enumerate(zip(src, tgt)):
                print(",".join([tokenizer.decode(s), tokenizer.decode(t)] + tokenizer.decode_batch(prediction_np[i])), file=f)

    def write_on_batch_end(self, trainer, pl_module, prediction, batch_indices, batch, batch_idx, dataloader_idx):
        self.write(pl_module.tgt_tokenizer, prediction, batch)
