"""RELION-style .star alignment files (interface of 1_Embedding/DM/Read_Alignment.py)."""
import numpy as np

STAR_COLUMNS = ['rlnAngleRot', 'rlnAngleTilt', 'rlnAnglePsi', 'rlnOriginX', 'rlnOriginY', 'rlnDefocusU',
                'rlnDefocusV', 'rlnVoltage', 'rlnSphericalAberration', 'rlnAmplitudeContrast', 'rlnDefocusAngle',
                'rlnCtfBfactor', 'rlnPhaseShift', 'rlnPixelSize', 'rlnImageName']


def parse_star(starfile):
    """pandas.DataFrame with one column per `_rln...` header label, one row per image
    (Read_Alignment.py:3-21: the header block ends at the first non-label line after a label)."""
    import pandas as pd
    headers, ln, found = [], 0, False
    with open(starfile, 'r') as f:
        for line in f:
            if line.startswith('_rln'):
                found = True
                headers.append(line.split('#')[0].rstrip().lstrip('_'))
            elif found:
                break
            ln += 1
    if not headers:
        raise ValueError('parse_star: no _rln header labels in %s' % starfile)
    star = pd.read_table(starfile, skiprows=ln, sep=r'\s+', header=None)
    star.columns = headers
    return star


def write_star(path, df, volt=300, Cs=2.7, ampc=0.1, pix=1.0, name='stack', angles=(0., 0., 0.)):
    """Alignment file in the layout of 0_Data_Inputs/Pristine_AddCtfSNR/GenStackAlign_CtfSnr.py:151-169,216-217."""
    df = np.asarray(df, dtype=float).reshape(-1)
    n = df.shape[0]
    col = lambda v: np.broadcast_to(np.asarray(v, dtype=float), (n,))
    volt, Cs, ampc, pix = col(volt), col(Cs), col(ampc), col(pix)
    with open(path, 'w') as f:
        f.write('\ndata_\n\nloop_\n\n')
        for i, c in enumerate(STAR_COLUMNS):
            f.write('_%s #%d\n' % (c, i + 1))
        for i in range(n):
            f.write('%.6f\t%.6f\t%.6f\t0\t0\t%r\t%r\t%r\t%r\t%r\t0\t0\t0\t%r\t%d@%s.mrcs\n'
                    % (angles[0], angles[1], angles[2], float(df[i]), float(df[i]), float(volt[i]), float(Cs[i]),
                       float(ampc[i]), float(pix[i]), i + 1, name))
