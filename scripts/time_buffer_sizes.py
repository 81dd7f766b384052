"""Latency of independent forwards at the buffer sizes of the reference's
scripts/time_buffer_sizes.py:13 (256 ... 32768 samples), CSV rows [model, device, buffer, seconds]."""
import click
import pandas as pd
import torch

from neural_waveshaping_synthesis_b200.timing import build_model, time_forward

BUFFER_SIZES = [256, 512, 1024, 2048, 4096, 8192, 16384, 32768]


@click.command()
@click.option("--gin-file", prompt="Model config gin file")
@click.option("--output-file", prompt="output file")
@click.option("--num-iters", default=100)
@click.option("--batch-size", default=1)
@click.option("--device", default="cuda:0")
@click.option("--length-in-seconds", default=4)
@click.option("--use-fast-newt", is_flag=True)
@click.option("--model-name", default="ours")
def main(gin_file, output_file, num_iters, batch_size, device, length_in_seconds, use_fast_newt, model_name):
    model = build_model(gin_file, use_fast_newt, device)
    rows = []
    with torch.no_grad():
        for bs in BUFFER_SIZES:
            frames = bs // 128
            control = torch.rand(batch_size, 2, frames, device=device)
            f0 = torch.rand(batch_size, 1, frames, device=device)
            for s in time_forward(lambda: model(f0, control), num_iters, device, warmup=10):
                rows.append([model_name, "cpu" if device == "cpu" else "gpu", bs, s])
    df = pd.DataFrame(rows)
    df.to_csv(output_file)
    print(df.groupby(2)[3].median().mul(1e3).round(4).to_string())


if __name__ == "__main__":
    main()
