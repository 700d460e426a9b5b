"""Known-answer table mined from the reference's own result logs (TEST INFRASTRUCTURE).

Sources (paths under the upstream repo):
  results/Set_dn_ADMM_L1/Set_dn_ADMM_L1.log:1156-1205   (S1 defaults 50 / 0.1 / 0.015)
  results/Set_dn_ADMM_CNC/Set_dn_ADMM_CNC.log:7110-7159 (S4 defaults 0.45 / 50 / 0.5 / 0.05 / 64)
  results/Set1_dn_ADMM_L1/Set1_dn_ADMM_L1.log:287-288, results/Set1_dn_ADMM_CNC/...log:399-400
Per-image PSNR (dB) for testsets/set/01..15.png; the L1 log prints 2 decimals,
the CNC log 4; the 4-decimal values below were regenerated with the unmodified
scripts in the build container and agree with every logged value.
"""

IMAGES = ['%02d' % i for i in range(1, 16)]
MASKS = ['Q_Random30', 'Q_Radial30', 'Q_Cartesian30']

L1_DEFAULTS = dict(iter_num=50, lambda1=0.1, reo=0.015)                       # S1:171
CNC_DEFAULTS = dict(alpha=0.45, iter_num=50, lambda1=0.5, reo=0.05, b=64)     # S4:176

PSNR = {
    ('Q_Random30', 'l1'): [24.3157, 23.6259, 23.8528, 23.7170, 23.8683, 23.5335, 24.1419, 26.0997,
                           23.9663, 26.1758, 26.3776, 23.4904, 23.2568, 23.1676, 24.2501],
    ('Q_Random30', 'cnc'): [24.7868, 24.3085, 24.7207, 24.4112, 24.5765, 24.2930, 24.6866, 25.8311,
                            24.5330, 25.8640, 26.1172, 24.3161, 24.1137, 24.0014, 24.9101],
    ('Q_Radial30', 'l1'): [24.2704, 23.4316, 23.7027, 23.3690, 23.6811, 23.5696, 24.1197, 26.0368,
                           24.5283, 26.1892, 26.0110, 23.3885, 23.0442, 23.0711, 24.1635],
    ('Q_Radial30', 'cnc'): [24.7046, 24.0266, 24.5205, 23.9644, 24.3232, 24.2582, 24.6197, 25.7543,
                            25.1068, 25.8342, 25.7511, 24.1609, 23.8157, 23.8371, 24.7534],
    ('Q_Cartesian30', 'l1'): [23.4156, 21.7765, 23.0237, 22.1211, 22.8470, 22.6791, 22.5838, 25.3767,
                              22.4229, 25.5496, 23.9576, 23.0593, 21.9595, 22.3445, 23.5009],
    ('Q_Cartesian30', 'cnc'): [23.8156, 22.2319, 23.7546, 22.6090, 23.4218, 23.2943, 22.9745, 25.2156,
                               22.8592, 25.3198, 23.8743, 23.8284, 22.6109, 23.0385, 24.0773],
}

# (avg PSNR, avg SSIM, avg RE) as printed by the logs (3 decimals)
AVERAGES = {
    ('Q_Random30', 'l1'): (24.256, 0.563, 0.198),
    ('Q_Random30', 'cnc'): (24.765, 0.496, 0.187),
    ('Q_Radial30', 'l1'): (24.172, 0.556, 0.200),
    ('Q_Radial30', 'cnc'): (24.629, 0.486, 0.190),
    ('Q_Cartesian30', 'l1'): (23.108, 0.513, 0.226),
    ('Q_Cartesian30', 'cnc'): (23.528, 0.446, 0.216),
}

# 05.png, Q_Random30 (BASELINE config 1): PSNR / SSIM / RE
SET1 = {'l1': (23.8683, 0.5877, 0.2028), 'cnc': (24.5765, 0.5600, 0.1870)}

# zero-filled PSNR printed at S1:101 for 05.png
ZERO_FILL_05 = {'Q_Random30': 20.6451, 'Q_Radial30': 20.8435, 'Q_Cartesian30': 20.3032}

DNCNN17_PARAMS = 555137   # "Params number: 555137", emitted at S6:431-432
