"""cbinfer_b200 -- the pycbinfer surface on a B200-native (sm_100a) backend.

Mirrors the reference's ``pycbinfer/__init__.py``: ``convert`` (:91-94), ``convertRecur``
(:20-45), ``subsitute`` (:10-17), ``mergeReLURecur`` (:47-66), ``propChangeIndexesOf1x1``
(:68-77), ``clearMemory`` (:79-82), ``getStateTensors`` (:84-89), ``tuneThresholdParameters``
(:98-144) and the classes ``CBConv2d`` / ``CBPoolMax2d``.  Same names, argument meaning and
return types; the per-frame work runs in ``libcbinfer_sm100.so`` (include/cbinfer_b200.h).
"""
import torch
import torch.nn as nn

from .conv2d import CBConv2d
from .conv2d import CBPoolMax2d
from .conv2d_cg import ChangeIndexes

__all__ = ['CBConv2d', 'CBPoolMax2d', 'ChangeIndexes', 'subsitute', 'convertRecur', 'mergeReLURecur',
           'propChangeIndexesOf1x1', 'clearMemory', 'getStateTensors', 'convert',
           'tuneThresholdParameters', 'convertPools', 'shareWorkspace']

verbose = False


def _log(fmt, *args):
    if verbose:
        print(fmt % args if args else fmt)


def _is(node, cls):
    """exact-type test, as the reference does (subclasses of nn.Conv2d / nn.ReLU are left alone)"""
    return type(node) is cls


def subsitute(node, threshold=1e-1, finegrained=False):
    """(node, replaced?) - an exact nn.Conv2d becomes a CBConv2d sharing its parameters
    (reference __init__.py:10-17; the misspelt name is the reference's)."""
    if not _is(node, nn.Conv2d):
        return node, False
    _log('replacing conv2d')
    cb = CBConv2d(node, threshold)
    cb.finegrained = finegrained
    return cb, True


def convertRecur(m, ignoreList=[], threshold=1e-1, finegrained=False):
    """One conversion pass over the children of `m` (reference __init__.py:20-45): containers
    recurse, Dropout and every type in `ignoreList` disappear, convolutions are substituted; the
    result is a fresh nn.Sequential with the surviving child names.  Returns (model, changed?)."""
    dropped = tuple(ignoreList) + (nn.Dropout,)
    out, changed = nn.Sequential(), False
    for name, child in m.named_children():
        if _is(child, nn.Sequential):
            # (like the reference, the recursion does not forward `finegrained`, __init__.py:28)
            child, hit = convertRecur(child, ignoreList, threshold)
        elif type(child) in dropped:
            _log('removing node %s', type(child))
            changed = True
            continue
        else:
            child, hit = subsitute(child, threshold=threshold, finegrained=finegrained)
        out.add_module(name, child)
        changed = changed or hit
    if changed:
        # the reference converts again until nothing changes (__init__.py:43-44); the second
        # pass only ever merges ReLUs
        out = convert(out, ignoreList)
    return out, changed


def mergeReLURecur(m):
    """Fold an nn.ReLU that directly follows a CBConv2d into that layer's `withReLU` flag and drop
    it from the container (reference __init__.py:47-66)."""
    kids = list(m.named_children())
    out = nn.Sequential()
    for pos, (name, child) in enumerate(kids):
        if _is(child, nn.Sequential):
            out.add_module(name, mergeReLURecur(child))
            continue
        if _is(child, CBConv2d):
            follower = kids[pos + 1][1] if pos + 1 < len(kids) else None
            if _is(follower, nn.ReLU):
                child.withReLU = True
        elif _is(child, nn.ReLU) and pos > 0 and _is(kids[pos - 1][1], CBConv2d):
            _log('merging ReLU layer')
            continue
        out.add_module(name, child)
    return out


def propChangeIndexesOf1x1(rootModule):
    """reference __init__.py:68-77.  Its test `m.kernel_size == [1,1]` compares a tuple with a list
    and therefore never holds; the behaviour (nothing is enabled) is kept on purpose."""
    for seq in [c for c in rootModule.modules() if _is(c, nn.Sequential)]:
        before = None
        for layer in seq:
            if _is(layer, CBConv2d) and _is(before, CBConv2d) and layer.kernel_size == [1, 1]:
                _log('enabling propagation of change indexes for 1x1')
                before.propChangeIndexes = True
            before = layer
    return rootModule


def _cb_layers(net):
    return [c for c in net.modules() if type(c) in (CBConv2d, CBPoolMax2d)]


def clearMemory(net):
    """drop the per-layer state of every CB layer (reference __init__.py:79-82)"""
    for layer in _cb_layers(net):
        layer.clearMemory()


def getStateTensors(net):
    """all state tensors of the CB layers in forward order (reference __init__.py:84-89)"""
    return [t for layer in _cb_layers(net) for t in layer.getStateTensors()]


def shareWorkspace(net):
    """Let all CBConv2d layers of one model instance share one stream-K workspace (their launches
    are ordered).  Models that may run concurrently must not share: call this per instance."""
    holder = None
    for m in net.modules():
        if type(m) == CBConv2d:
            if holder is None:
                holder = m._workspace_holder()
            m._wsHolder = holder
    return net


def convert(m, ignoreList=[], threshold=1e-1):
    """nn.Conv2d -> CBConv2d, Dropout removed, ReLUs folded (reference __init__.py:91-94); returns an
    nn.Sequential with the original child names."""
    converted, _ = convertRecur(m, ignoreList=ignoreList, threshold=threshold)
    return shareWorkspace(mergeReLURecur(converted))


def convertPools(m):
    """Helper for the hand-written recipe of the reference (sceneLabeling/modelLoader.py:72-78):
    wrap every 2x2/stride-2 nn.MaxPool2d that directly follows a CBConv2d in CBPoolMax2d and let
    that conv propagate its change indexes.  Not part of the reference API (it does this by
    hand); returns a new nn.Sequential with the same child names."""
    mout = nn.Sequential()
    prev = None
    for name, node in m.named_children():
        if type(node) is nn.Sequential:
            node = convertPools(node)
        elif type(node) is nn.MaxPool2d and type(prev) is CBConv2d and not prev.finegrained:
            ks = node.kernel_size if isinstance(node.kernel_size, tuple) else (node.kernel_size,) * 2
            st = node.stride if isinstance(node.stride, tuple) else (node.stride,) * 2
            if ks == (2, 2) and st == (2, 2) and node.padding in (0, (0, 0)):
                prev.propChangeIndexes = True
                node = CBPoolMax2d(node)
        mout.add_module(name, node)
        prev = node
    return mout


def tuneThresholdParameters(vidSeqReader, evalSequences, numFramesPerSeq,
                            targetGenerator, preprocessor,
                            modelBaseline, modelTest, evaluator,
                            cbModuleList, lossToleranceList, initThreshold=1e-2, thresholdIncrFactor=1.2):
    """Greedy front-to-back threshold search (reference __init__.py:98-144).  Layer by layer: start
    at `initThreshold`, multiply by `thresholdIncrFactor` for as long as the summed loss over the
    evaluation sequences stays within that layer's tolerance of the loss before the layer was
    touched, then undo the last step.  A scalar tolerance applies to every layer."""
    tolerances = lossToleranceList if type(lossToleranceList) is list \
        else [lossToleranceList] * len(cbModuleList)
    assert len(tolerances) == len(cbModuleList)

    def summed_loss():
        clearMemory(modelTest)
        total = 0
        with torch.no_grad():
            for seq in evalSequences:
                frames, target = vidSeqReader.getDataFrames(seqName=seq, numFrames=numFramesPerSeq)
                if target is None:
                    target = targetGenerator(frames[-1])
                out = None
                for frame in frames:
                    out = modelTest(preprocessor(frame).cuda())
                total += evaluator(out, target)
        return total

    reference_loss = summed_loss()
    for pos, layer in enumerate(cbModuleList):
        _log('adjusting threshold for module %d of %d', pos + 1, len(cbModuleList))
        layer.threshold = initThreshold
        while True:
            layer.threshold *= thresholdIncrFactor
            loss = summed_loss()
            _log('. (%f < %f + %f)', loss, reference_loss, tolerances[pos])
            if loss - reference_loss > tolerances[pos]:
                layer.threshold /= thresholdIncrFactor
                break
        reference_loss = summed_loss()
    return modelTest
