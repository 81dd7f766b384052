"""Real-time factor of one forward pass — same options as the reference's
scripts/time_forward_pass.py:14-22, timed with CUDA events (device default cuda:0: this
implementation has no CPU path)."""
import click
import numpy as np
import torch

from neural_waveshaping_synthesis_b200.timing import build_model, time_forward


@click.command()
@click.option("--gin-file", prompt="Model config gin file")
@click.option("--num-iters", default=100)
@click.option("--batch-size", default=1)
@click.option("--device", default="cuda:0")
@click.option("--length-in-seconds", default=4)
@click.option("--sample-rate", default=16000)
@click.option("--control-hop", default=128)
@click.option("--use-fast-newt", is_flag=True)
def main(gin_file, num_iters, batch_size, device, length_in_seconds, sample_rate, control_hop, use_fast_newt):
    frames = sample_rate * length_in_seconds // control_hop
    model = build_model(gin_file, use_fast_newt, device)
    control = torch.rand(batch_size, 2, frames, device=device)
    f0 = torch.rand(batch_size, 1, frames, device=device)
    with torch.no_grad():
        secs = np.array(time_forward(lambda: model(f0, control), num_iters, device))
    rtf = secs / length_in_seconds
    print("forward: mean %.3f ms  min %.3f  max %.3f  (n=%d, batch %d)" %
          (secs.mean() * 1e3, secs.min() * 1e3, secs.max() * 1e3, num_iters, batch_size))
    print("Mean RTF: %.6f" % rtf.mean())
    print("90th percentile RTF: %.6f" % np.percentile(rtf, 90))
    print("samples/s: %.4g" % (batch_size * frames * control_hop / secs.mean()))


if __name__ == "__main__":
    main()
