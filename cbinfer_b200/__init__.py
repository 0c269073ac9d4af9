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
           'tuneThresholdParameters', 'convertPools']

verbose = False


def _log(msg):
    if verbose:
        print(msg)


def subsitute(node, threshold=1e-1, finegrained=False):
    if type(node) is torch.nn.modules.conv.Conv2d:
        _log('replacing conv2d')
        m = CBConv2d(node, threshold)
        m.finegrained = finegrained
        return m, True
    else:
        return node, False


def convertRecur(m, ignoreList=[], threshold=1e-1, finegrained=False):
    changed = False

    mout = nn.Sequential()
    for i, (nodeName, node) in enumerate(m.named_children()):
        if type(node) in [nn.Sequential]:
            # handle nn.Sequential containers through recursion
            # (the reference drops `finegrained` here, __init__.py:28; kept for parity)
            msub, c = convertRecur(node, ignoreList, threshold)
            mout.add_module(nodeName, msub)
            changed |= c
        elif type(node) in ignoreList + [nn.Dropout]:
            # remove nodes not needed during inference (e.g. Dropout, and those in the ignoreList)
            _log('removing node %s' % (type(node),))
            changed = True
            continue
        else:
            # handle simple substitutions (i.e. convert Conv2d to CBconv2d)
            nodeOut, newNode = subsitute(node, threshold=threshold, finegrained=finegrained)
            mout.add_module(nodeName, nodeOut)
            changed |= newNode

    # another round until convergence (a no-op apart from the ReLU merge, __init__.py:43-44)
    if changed:
        mout = convert(mout, ignoreList)
    return mout, changed


def mergeReLURecur(m):
    mout = nn.Sequential()
    for i, (nodeName, node) in enumerate(m.named_children()):
        # handle nn.Sequential containers through recursion
        if type(node) in [nn.Sequential]:
            mout.add_module(nodeName, mergeReLURecur(node))
            continue
        # enable built-in ReLU of CBconv
        elif type(node) in [CBConv2d]:
            chldrn = list(m.children())
            if len(chldrn) > i + 1 and type(chldrn[i + 1]) is torch.nn.modules.activation.ReLU:
                node.withReLU = True
        # remove ReLU if CBconv layer proceeded
        elif type(node) is torch.nn.modules.activation.ReLU and i >= 1 and type(list(m.children())[i - 1]) is CBConv2d:
            _log('merging ReLU layer')
            continue  # i.e. don't add the module!!

        mout.add_module(nodeName, node)
    return mout


def propChangeIndexesOf1x1(rootModule):
    seqContainers = list(filter(lambda m: type(m) == torch.nn.Sequential, rootModule.modules()))
    for seqCont in seqContainers:
        mPrev = None
        for m in seqCont:
            # the reference compares kernel_size (a tuple) with the list [1,1] (__init__.py:73), so
            # this branch never fires there; kept verbatim in behaviour.
            if type(m) == CBConv2d and type(mPrev) == CBConv2d and m.kernel_size == [1, 1]:
                _log('enabling propagation of change indexes for 1x1')
                mPrev.propChangeIndexes = True
            mPrev = m
    return rootModule


def clearMemory(net):
    for m in net.modules():
        if type(m) == CBConv2d or type(m) == CBPoolMax2d:
            m.clearMemory()


def getStateTensors(net):
    state = []
    for m in net.modules():
        if type(m) == CBConv2d or type(m) == CBPoolMax2d:
            state += m.getStateTensors()
    return state


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
    m1, changed = convertRecur(m, ignoreList=ignoreList, threshold=threshold)
    mout = mergeReLURecur(m1)
    return shareWorkspace(mout)


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
    """Greedy front-to-back threshold search (reference __init__.py:98-144): for each CB module
    raise its threshold by `thresholdIncrFactor` while the loss increase stays within the module's
    tolerance, then step back once."""
    if type(lossToleranceList) is not list:
        # if loss tolerance is given as a single value, apply it to all modules
        lossToleranceList = [lossToleranceList] * len(cbModuleList)
    assert(len(cbModuleList) == len(lossToleranceList))

    def evaluateModel():
        clearMemory(modelTest)
        totalLoss = 0
        with torch.no_grad():
            for seqName in evalSequences:
                frames, target = vidSeqReader.getDataFrames(seqName=seqName, numFrames=numFramesPerSeq)
                if target is None:
                    target = targetGenerator(frames[-1])
                for frame in frames:
                    feedData = preprocessor(frame)
                    outTest = modelTest(feedData.cuda())
                loss = evaluator(outTest, target)
                totalLoss += loss
        return totalLoss

    # greedy front-to-back threshold adjustment
    prevLoss = evaluateModel()
    for i, m in enumerate(cbModuleList):
        _log('adjusting threshold for module %d of %d' % (i + 1, len(cbModuleList)))
        m.threshold = initThreshold  # initialize threshold value
        while True:
            # increase th while loss ok. Once insufficient, take 1 step back.
            m.threshold *= thresholdIncrFactor
            loss = evaluateModel()
            _log('. (%f < %f + %f)' % (loss, prevLoss, lossToleranceList[i],))
            if loss - prevLoss > lossToleranceList[i]:
                m.threshold /= thresholdIncrFactor
                break
        prevLoss = evaluateModel()
    return modelTest
