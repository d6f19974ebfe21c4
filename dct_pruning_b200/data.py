"""Real-data ingestion for the scoring pass: the loaders `imp_score` feeds from.

    load_data             /root/reference/utils/common.py:57-161   (cifar10 :60-75, imagenet :77-121, DUTS :123-159)
    SalObjDataset & co.   /root/reference/data/data_loader.py:15-268   (RescaleT :15-45, RandomCrop :80-108, ToTensorLab flag 0 :153-246)

Same datasets, directory layout, transforms and batch sizes as the reference's *training* loaders (the ones `imp_score`
iterates, common.py:374).  Two deliberate differences, both opt-in-free and result-neutral for a fixed seed:
  * sampling is seeded (`seed`): the reference reshuffles unseeded for every hook site, so no two of its runs agree;
    one seeded pass over `limit` batches feeds every site here;
  * the DUTS pipeline decodes with PIL and resizes with torch's bilinear kernel (skimage is not a dependency); value
    range, crop size (288 out of 320, common.py:154-155), flip rule and normalisation are the reference's.
Under torchrun every rank builds the same loader with the same seed and takes its own slice of each batch
(generate.inference), so the union of the shards is exactly the single-process batch.
"""
import glob
import os
import random

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

CIFAR_MEAN, CIFAR_STD = (0.4914, 0.4822, 0.4465), (0.2023, 0.1994, 0.2010)
IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


class SalObjImages(Dataset):
    """Images of DUTS-TR through RescaleT(320) -> RandomCrop(288) -> ToTensorLab(flag=0); labels are not needed for scoring
    but the sample keeps the reference's {'imidx', 'image', 'label'} shape (label all zeros)."""

    def __init__(self, img_name_list, rescale=320, crop=288):
        self.image_name_list = list(img_name_list)
        self.rescale, self.crop = rescale, crop

    def __len__(self):
        return len(self.image_name_list)

    def __getitem__(self, idx):
        from PIL import Image
        with Image.open(self.image_name_list[idx]) as im:
            image = np.asarray(im.convert('RGB') if im.mode not in ('L', 'RGB') else im, dtype=np.float64) / 255.0
        if image.ndim == 2:
            image = image[:, :, None]
        t = torch.from_numpy(image).permute(2, 0, 1)[None]                         # RescaleT: square resize, [0,1] floats
        t = torch.nn.functional.interpolate(t, size=(self.rescale, self.rescale), mode='bilinear', align_corners=False, antialias=True)[0]
        if random.random() >= 0.5:                                                 # RandomCrop: vertical flip, then the crop
            t = t.flip(1)
        top = np.random.randint(0, self.rescale - self.crop)
        left = np.random.randint(0, self.rescale - self.crop)
        t = t[:, top:top + self.crop, left:left + self.crop]
        t = t / t.max().clamp_min(1e-12)                                           # ToTensorLab(flag=0)
        if t.shape[0] == 1:
            t = ((t - 0.485) / 0.229).expand(3, -1, -1).clone()
        else:
            mean = torch.tensor(IMAGENET_MEAN, dtype=t.dtype).view(3, 1, 1)
            std = torch.tensor(IMAGENET_STD, dtype=t.dtype).view(3, 1, 1)
            t = (t - mean) / std
        return {'imidx': torch.tensor([idx]), 'image': t, 'label': torch.zeros(1, self.crop, self.crop, dtype=t.dtype)}


def _seed_worker(worker_id):
    s = torch.initial_seed() % (1 << 31)
    np.random.seed(s)
    random.seed(s)


def load_data(args, seed=0):
    """The training loader `imp_score` iterates for args.dataset (cifar10 | imagenet | DUTS), plus the validation loader
    where the reference builds one.  Raises FileNotFoundError with the expected layout when the data are absent (the
    reference would try to download CIFAR-10; this build never touches the network)."""
    from torchvision import datasets, transforms
    gen = torch.Generator().manual_seed(seed)
    kw = dict(generator=gen, worker_init_fn=_seed_worker)
    np.random.seed(seed)                                 # in-process transforms (num_workers = 0) draw from the global generators
    random.seed(seed)
    torch.manual_seed(seed)
    if args.dataset == 'cifar10':
        if not os.path.isdir(os.path.join(args.data_dir, 'cifar-10-batches-py')):
            raise FileNotFoundError('%s/cifar-10-batches-py not found (CIFAR-10 python batches; nothing is downloaded)' % args.data_dir)
        train_tf = transforms.Compose([transforms.RandomCrop(32, padding=4), transforms.RandomHorizontalFlip(), transforms.ToTensor(),
                                       transforms.Normalize(CIFAR_MEAN, CIFAR_STD)])
        test_tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(CIFAR_MEAN, CIFAR_STD)])
        train = datasets.CIFAR10(root=args.data_dir, train=True, download=False, transform=train_tf)
        test = datasets.CIFAR10(root=args.data_dir, train=False, download=False, transform=test_tf)
        return (DataLoader(train, batch_size=args.batch_size, shuffle=True, num_workers=1, **kw),
                DataLoader(test, batch_size=args.batch_size, shuffle=False, num_workers=1))
    if args.dataset == 'imagenet':
        traindir, valdir = os.path.join(args.data_dir, 'ILSVRC2012_img_train'), os.path.join(args.data_dir, 'val')
        if not os.path.isdir(traindir):
            raise FileNotFoundError('%s not found (ImageFolder layout: <data_dir>/ILSVRC2012_img_train/<class>/*.JPEG)' % traindir)
        norm = transforms.Normalize(mean=IMAGENET_MEAN, std=IMAGENET_STD)
        train = datasets.ImageFolder(traindir, transforms.Compose([transforms.RandomResizedCrop(224), transforms.RandomHorizontalFlip(),
                                                                   transforms.Resize(224), transforms.ToTensor(), norm]))
        val_loader = None
        if os.path.isdir(valdir):
            val = datasets.ImageFolder(valdir, transforms.Compose([transforms.Resize(256), transforms.CenterCrop(224), transforms.Resize(224),
                                                                   transforms.ToTensor(), norm]))
            val_loader = DataLoader(val, batch_size=args.batch_size, shuffle=False, num_workers=8, pin_memory=True)
        workers = int(getattr(args, 'workers', 8))
        return DataLoader(train, batch_size=args.batch_size, shuffle=True, num_workers=workers, pin_memory=True, **kw), val_loader
    if args.dataset == 'DUTS':
        image_dir = os.path.join(args.data_dir, 'DUTS-TR', 'DUTS-TR-Image')
        names = sorted(glob.glob(os.path.join(image_dir, '*.jpg')))
        if not names:
            raise FileNotFoundError('%s/*.jpg not found (DUTS-TR images)' % image_dir)
        print('---\ntrain images: ', len(names), '\n---')
        return DataLoader(SalObjImages(names), batch_size=args.batch_size, shuffle=True, num_workers=1, **kw), None
    raise ValueError('unknown dataset %r' % (args.dataset,))
