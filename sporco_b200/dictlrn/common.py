"""Field and header maps of the dictionary-learning record for an ADMM or PGM X step and a PGM or
ADMM (consensus) D step (mirror of sporco/dictlrn/common.py:19-135)."""


def evlmap(accdfid):
    return {'ObjFun': 'ObjFun', 'DFid': 'DFid', 'RegL1': 'RegL1'} if accdfid else {}


def isxmap(xmethod, opt):
    if xmethod == 'admm':
        isx = {'XPrRsdl': 'PrimalRsdl', 'XDlRsdl': 'DualRsdl', 'XRho': 'Rho'}
    else:
        isx = {'X_F_Btrack': 'F_Btrack', 'X_Q_Btrack': 'Q_Btrack', 'X_ItBt': 'IterBTrack',
               'X_L': 'L', 'X_Rsdl': 'Rsdl'}
    if not opt['AccurateDFid']:
        isx.update(evlmap(True))
    return isx


def isdmap(dmethod):
    if dmethod == 'pgm':
        return {'Cnstr': 'Cnstr', 'D_F_Btrack': 'F_Btrack', 'D_Q_Btrack': 'Q_Btrack',
                'D_ItBt': 'IterBTrack', 'D_L': 'L', 'D_Rsdl': 'Rsdl'}
    return {'Cnstr': 'Cnstr', 'DPrRsdl': 'PrimalRsdl', 'DDlRsdl': 'DualRsdl', 'DRho': 'Rho'}


def _xcols(xmethod, opt):
    if xmethod == 'admm':
        return [('r_X', 'XPrRsdl'), ('s_X', 'XDlRsdl'), (u'ρ_X', 'XRho')], []
    if opt['CBPDN', 'Backtrack'] is not None:
        return [('F_X', 'X_F_Btrack'), ('Q_X', 'X_Q_Btrack'), ('It_X', 'X_ItBt'),
                ('L_X', 'X_L')], ['X_Rsdl']
    return [('L_X', 'X_L')], ['X_Rsdl']


def _dcols(opt, dmethod='pgm'):
    if dmethod != 'pgm':
        return [('r_D', 'DPrRsdl'), ('s_D', 'DDlRsdl'), (u'ρ_D', 'DRho')], []
    if opt['CCMOD', 'Backtrack'] is not None:
        return [('F_D', 'D_F_Btrack'), ('Q_D', 'D_Q_Btrack'), ('It_D', 'D_ItBt'),
                ('L_D', 'D_L')], ['D_Rsdl']
    return [('L_D', 'D_L')], ['D_Rsdl']


def isfld(xmethod, dmethod, opt):
    fld = ['Iter', 'ObjFun', 'DFid', 'RegL1', 'Cnstr']
    for cols, extra in (_xcols(xmethod, opt), _dcols(opt, dmethod)):
        fld.extend([f for _, f in cols] + extra)
    fld.append('Time')
    return fld


def hdrtxt(xmethod, dmethod, opt):
    txt = ['Itn', 'Fnc', 'DFid', u'ℓ1', 'Cnstr']
    for cols, _ in (_xcols(xmethod, opt), _dcols(opt, dmethod)):
        txt.extend([h for h, _ in cols])
    return txt


def hdrmap(xmethod, dmethod, opt):
    hdr = {'Itn': 'Iter', 'Fnc': 'ObjFun', 'DFid': 'DFid', u'ℓ1': 'RegL1', 'Cnstr': 'Cnstr'}
    for cols, _ in (_xcols(xmethod, opt), _dcols(opt, dmethod)):
        hdr.update(dict(cols))
    return hdr
