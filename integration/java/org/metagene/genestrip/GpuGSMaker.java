// GpuGSMaker.java -- installs the GPU matcher and the GPU filter through the reference's own factory methods, the way the
// reference's ComprehensiveFilterTest swaps goals (core/src/test/.../goals/refseq/ComprehensiveFilterTest.java:91-152):
//   createGoalChainForMatchResult (core/.../GSMaker.java:560-583)  -> MatchResultGoal whose createMatcher (:174-197 of
//                                                                     goals/MatchResultGoal.java) returns GpuFastqKMerMatcher
//   createGoalChainForFilter      (:602-625)                        -> GpuFilterGoal
//   createExecutionContext        (:90-93)                          -> zero consumer threads: the parser thread fills the
//                                                                     pinned batches inline, the GPU is the consumer
// Everything else (db goals, CSV writers, CLI) is the reference's, untouched.  `new GpuGSMaker<>(project).match(...)` /
// `.filter(...)`, or return it from Main.createMaker (core/.../Main.java:143-145).
package org.metagene.genestrip;

import java.util.Map;

import org.metagene.genestrip.goals.FastqDownloadsGoal;
import org.metagene.genestrip.goals.FastqMapGoal;
import org.metagene.genestrip.goals.FastqMapTransformGoal;
import org.metagene.genestrip.goals.FilterGoal;
import org.metagene.genestrip.goals.GpuFilterGoal;
import org.metagene.genestrip.goals.LoadDBGoal;
import org.metagene.genestrip.goals.LoadIndexGoal;
import org.metagene.genestrip.goals.MatchResultGoal;
import org.metagene.genestrip.gpu.GsNative;
import org.metagene.genestrip.io.StreamingResourceStream;
import org.metagene.genestrip.make.ObjectGoal;
import org.metagene.genestrip.match.FastqKMerMatcher;
import org.metagene.genestrip.match.GpuFastqKMerMatcher;
import org.metagene.genestrip.store.KMerStore;
import org.metagene.genestrip.store.TunableKMerStore;
import org.metagene.genestrip.tax.SmallTaxTree;
import org.metagene.genestrip.tax.SmallTaxTree.SmallTaxIdNode;

public class GpuGSMaker<P extends GSProject> extends GSMaker<P> {
    private final long ctx;
    private long deviceDb;                 // one upload per maker: every key of the run matches against the same device database
    private KMerStore<SmallTaxIdNode> uploaded;
    private GpuFilterGoal<P> filterGoal;

    public GpuGSMaker(P project, int... devices) {
        super(project);
        ctx = GsNative.ctxCreate(devices.length == 0 ? null : devices);
    }

    @Override
    protected ExecutionContext createExecutionContext(Thread mainThread, P project) {
        return new DefaultExecutionContext(mainThread, 0, project.longConfigValue(GSConfigKey.LOG_PROGRESS_UPDATE_CYCLE));
    }

    @Override
    protected MatchResultGoal<P> createGoalChainForMatchResult(boolean lr, String key, String... pathsOrURLs) {
        ObjectGoal<Map<String, StreamingResourceStream>, P> fastqMapGoal = new FastqMapGoal<P>(getProject(), true, getGoal(GSGoalKey.SETUP)) {
            @Override
            protected void doMakeThis() {
                set(createFastqMap(key, pathsOrURLs, null, null, null));
            }
        };
        ObjectGoal<Map<String, StreamingResourceStream>, P> transformed = new FastqMapTransformGoal<P>(getProject(), true, fastqMapGoal,
                getGoal(GSGoalKey.SETUP));
        FastqDownloadsGoal<P> downloads = new FastqDownloadsGoal<P>(getProject(), true, fastqMapGoal, transformed, getGoal(GSGoalKey.SETUP));
        LoadDBGoal<P> loadDBGoal = getLoadDBGoal();
        return new MatchResultGoal<P>(getProject(), lr ? GSGoalKey.MATCHRESLR : GSGoalKey.MATCHRES, transformed, loadDBGoal,
                getExecutionContext(getProject()), getGoal(GSGoalKey.SETUP), downloads) {
            @Override
            protected FastqKMerMatcher createMatcher(KMerStore<SmallTaxIdNode> store, SmallTaxTree taxTree, ExecutionContext bundle, boolean withProbs,
                    String dbMD5) {
                if (uploaded != store) {
                    if (deviceDb != 0) {
                        GsNative.dbDestroy(deviceDb);
                    }
                    deviceDb = GpuFastqKMerMatcher.upload(ctx, store);
                    uploaded = store;
                }
                boolean useBloom = !(store instanceof TunableKMerStore) || ((TunableKMerStore<SmallTaxIdNode>) store).isUseFilter();
                return new GpuFastqKMerMatcher(deviceDb, useBloom, 1 << 20, 256L << 20, store, intConfigValue(GSConfigKey.INITIAL_READ_SIZE_BYTES),
                        intConfigValue(GSConfigKey.THREAD_QUEUE_SIZE), bundle, withProbs, intConfigValue(GSConfigKey.MAX_KMER_RES_COUNTS), taxTree,
                        intConfigValue(GSConfigKey.MAX_CLASSIFICATION_PATHS), doubleConfigValue(GSConfigKey.MAX_READ_TAX_ERROR_COUNT),
                        doubleConfigValue(GSConfigKey.MAX_READ_CLASS_ERROR_COUNT), booleanConfigValue(GSConfigKey.WRITE_ALL),
                        intConfigValue(GSConfigKey.MIN_KMERS_FOR_CLASS), dbMD5) {
                    @Override
                    protected boolean isProgressBar() {
                        return booleanConfigValue(GSConfigKey.PROGRESS_BAR);
                    }

                    @Override
                    protected String getProgressBarTaskName() {
                        return getKey().getName();
                    }
                };
            }
        };
    }

    @Override
    @SuppressWarnings("unchecked")
    protected FilterGoal<P> createGoalChainForFilter(String key, String... pathsOrURLs) {
        ObjectGoal<Map<String, StreamingResourceStream>, P> fastqMapGoal = new FastqMapGoal<P>(getProject(), true, getGoal(GSGoalKey.SETUP)) {
            @Override
            protected void doMakeThis() {
                set(createFastqMap(key, pathsOrURLs, null, null, null));
            }
        };
        ObjectGoal<Map<String, StreamingResourceStream>, P> transformed = new FastqMapTransformGoal<P>(getProject(), true, fastqMapGoal,
                getGoal(GSGoalKey.SETUP));
        FastqDownloadsGoal<P> downloads = new FastqDownloadsGoal<P>(getProject(), true, fastqMapGoal, transformed, getGoal(GSGoalKey.SETUP));
        LoadIndexGoal<P> index = (LoadIndexGoal<P>) getGoal(GSGoalKey.LOAD_INDEX);
        if (filterGoal != null) {
            filterGoal.releaseDevice();
        }
        filterGoal = new GpuFilterGoal<P>(ctx, getProject(), transformed, index, getExecutionContext(getProject()), getGoal(GSGoalKey.SETUP), downloads);
        return filterGoal;
    }

    @Override
    public void dumpAll() {
        super.dumpAll();
        if (filterGoal != null) {
            filterGoal.releaseDevice();
            filterGoal = null;
        }
        if (deviceDb != 0) {
            GsNative.dbDestroy(deviceDb);
            deviceDb = 0;
            uploaded = null;
        }
    }
}
