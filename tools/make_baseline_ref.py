"""Stage a runnable copy of the UNMODIFIED reference under the git-ignored `baseline/_ref/` so that it travels to the GPU box with the
gpurun snapshot (`/root/reference` does not exist there).  Nothing under baseline/_ref is ever committed or
imported by the product; it is the denominator of the "x times the reference single-GPU PyTorch" figure (tools/ref_gpu_timing.py) and,
when present, the CPU arm of `bench.py --impl reference` (cpu_baseline.kind = "reference").

    python tools/make_baseline_ref.py            # copies model/ utils/ configs/ dataset/transforms.py scripts/train_utils.py example.py

Only Python sources and YAML configs are copied (no checkpoints, no media: the S3D weights and data/ stay behind).  The three shims the
reference needs in this image (omegaconf, timm, transformers-4.27 names) stay in tests/golden/_ref_import.py and are installed in memory."""
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('SYNCHFORMER_REF_SRC', '/root/reference')
DST = os.path.join(REPO, 'baseline', '_ref')
KEEP_EXT = ('.py', '.yaml', '.yml')


def main():
    if not os.path.isdir(os.path.join(SRC, 'model')):
        print(f'{SRC} not found: nothing staged (on the GPU box baseline/_ref must already be in the snapshot)')
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for top in ('model', 'utils', 'configs', 'dataset', 'scripts'):
        for root, _, files in os.walk(os.path.join(SRC, top)):
            for f in files:
                if f.endswith(KEEP_EXT):
                    rel = os.path.relpath(os.path.join(root, f), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copyfile(os.path.join(SRC, rel), os.path.join(DST, rel))
                    n += 1
    for f in ('example.py', 'main.py', 'LICENSE'):
        if os.path.exists(os.path.join(SRC, f)):
            shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
            n += 1
    print(f'staged {n} files of the unmodified reference under {DST} (git-ignored)')
    return 0


if __name__ == '__main__':
    sys.exit(main())
