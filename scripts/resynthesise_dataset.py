"""Resynthesises a dataset split with a trained checkpoint — same options as the reference's
scripts/resynthesise_dataset.py:14-25.  Batches are staged through pinned memory and the D2H copy /
wav writing of batch i overlaps the forward of batch i+1."""
import os

import click
import torch
from scipy.io import wavfile
from tqdm import tqdm

from neural_waveshaping_synthesis.data.urmp import URMPDataset
from neural_waveshaping_synthesis.utils import make_dir_if_not_exists
from neural_waveshaping_synthesis_b200.timing import build_model


def _write(output_path, sample_rate, names, target, output):
    for name, tgt, out in zip(names, target, output):
        wavfile.write(os.path.join(output_path, "%s.target.wav" % name), sample_rate, tgt)
        wavfile.write(os.path.join(output_path, "%s.output.wav" % name), sample_rate, out)


@click.command()
@click.option("--model-gin", prompt="Model .gin file")
@click.option("--model-checkpoint", prompt="Model checkpoint")
@click.option("--dataset-root", prompt="Dataset root directory")
@click.option("--dataset-split", default="test")
@click.option("--output-path", default="audio_output")
@click.option("--load-data-to-memory", default=False)
@click.option("--device", default="cuda:0")
@click.option("--batch-size", default=8)
@click.option("--num_workers", default=16)
@click.option("--use-fastnewt", is_flag=True)
def main(model_gin, model_checkpoint, dataset_root, dataset_split, output_path, load_data_to_memory, device,
         batch_size, num_workers, use_fastnewt):
    make_dir_if_not_exists(output_path)
    data = URMPDataset(dataset_root, dataset_split, load_data_to_memory)
    loader = torch.utils.data.DataLoader(data, batch_size=batch_size, num_workers=num_workers, pin_memory=True)
    model = build_model(model_gin, use_fastnewt, torch.device(device), checkpoint=model_checkpoint)
    sample_rate = int(model.sample_rate)
    from neural_waveshaping_synthesis_b200.streaming import HostPipeline

    def batches():
        for batch in tqdm(loader):
            yield batch["f0"].float(), batch["control"].float(), (batch["name"], batch["audio"].float().numpy())

    # upload of the next batch and download of the previous result overlap the forward of the current one
    for (names, target), audio in HostPipeline(model, torch.device(device)).run(batches()):
        _write(output_path, sample_rate, names, target, audio.numpy())


if __name__ == "__main__":
    main()
